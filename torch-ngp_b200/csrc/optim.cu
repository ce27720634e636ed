// optim.cu — fused mixed-precision optimizer step for the hash table and MLP weights.
//
// Replaces, for the parameters of the hot path, the reference trainer's sequence (nerf/utils.py:866-868 +
// main_nerf.py:132): GradScaler.unscale_ (read+write every grad), the inf/NaN check, torch.optim.Adam(betas=(0.9,
// 0.99), eps=1e-15) (4 reads + 3 writes per parameter), the fp32->fp16 table cast of the next forward (grid.py:43-44)
// and the gradient zeroing — one pass: read fp16 grad (as produced by the scatter kernel / allreduce), unscale, Adam
// update of fp32 master + moments, write the fp16 shadow, zero the grad.  A skipped step (non-finite gradient, as
// GradScaler does) leaves parameters and moments untouched; scale growth/backoff follows torch.amp.GradScaler
// (growth 2.0 every 2000 clean steps, backoff 0.5) and lives on the device: no host synchronisation.
#include "common.cuh"

namespace ngp {

struct ScalerState {      // device-resident GradScaler state
    float scale;          // current loss scale
    int growth_tracker;   // clean steps since the last change
    int found_inf;        // set by k_check_finite for the current step
    int step;             // number of optimizer steps actually taken (bias correction)
    float lr_scale;       // multiplies the lr argument (LambdaLR-style schedules without re-capturing a CUDA graph)
    int reserved[3];
};

template <typename G>
__device__ __forceinline__ float gload(const G* g, size_t i);
template <> __device__ __forceinline__ float gload<__half>(const __half* g, size_t i) { return __half2float(g[i]); }
template <> __device__ __forceinline__ float gload<float>(const float* g, size_t i) { return g[i]; }

template <typename G>
__global__ void k_check_finite(const G* __restrict__ grads, size_t n, ScalerState* __restrict__ st) {
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = gload<G>(grads, i);
        bad |= !isfinite(v);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31u) == 0) atomicOr(&st->found_inf, 1);
}

template <typename G>
__global__ void k_adam_step(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, G* __restrict__ g,
                            __half* __restrict__ shadow, size_t n, float lr, float beta1, float beta2, float eps,
                            const ScalerState* __restrict__ st, int zero_grad) {
    if (st->found_inf) {
        // skipped step: only clear the gradient for the next accumulation
        if (zero_grad)
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) g[i] = (G)0;
        return;
    }
    const float inv_scale = 1.0f / st->scale;
    const int step = st->step + 1;
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2 = 1.0f - powf(beta2, (float)step);
    const float step_size = lr * st->lr_scale / bc1;
    const float rsqrt_bc2 = rsqrtf(bc2);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = gload<G>(g, i) * inv_scale;
        const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gi);
        const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gi * gi);
        const float denom = sqrtf(vi) * rsqrt_bc2 + eps;        // torch.optim.Adam: sqrt(v)/sqrt(bc2) + eps
        const float pi = p[i] - step_size * (mi / denom);
        m[i] = mi; v[i] = vi; p[i] = pi;
        if (shadow) shadow[i] = __float2half_rn(pi);
        if (zero_grad) g[i] = (G)0;
    }
}

// 4 parameters per thread per iteration: 128-bit loads/stores of p/m/v, 64-bit of the fp16 grad and shadow
__global__ void k_adam_step_vec4(float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v, uint2* __restrict__ g,
                                 uint2* __restrict__ shadow, size_t n4, float lr, float beta1, float beta2, float eps,
                                 const ScalerState* __restrict__ st, int zero_grad) {
    if (st->found_inf) {
        if (zero_grad)
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) g[i] = make_uint2(0, 0);
        return;
    }
    const float inv_scale = 1.0f / st->scale;
    const int step = st->step + 1;
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2 = 1.0f - powf(beta2, (float)step);
    const float step_size = lr * st->lr_scale / bc1;
    const float rsqrt_bc2 = rsqrtf(bc2);
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint2 gr = g[i];
        const float2 g01 = __half22float2(*reinterpret_cast<const __half2*>(&gr.x));
        const float2 g23 = __half22float2(*reinterpret_cast<const __half2*>(&gr.y));
        const float gi[4] = {g01.x * inv_scale, g01.y * inv_scale, g23.x * inv_scale, g23.y * inv_scale};
        float4 pp = p[i], mm = m[i], vv = v[i];
        float* pf = reinterpret_cast<float*>(&pp); float* mf = reinterpret_cast<float*>(&mm); float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            mf[k] = fmaf(beta1, mf[k], omb1 * gi[k]);
            vf[k] = fmaf(beta2, vf[k], omb2 * gi[k] * gi[k]);
            pf[k] = pf[k] - step_size * (mf[k] / (sqrtf(vf[k]) * rsqrt_bc2 + eps));
        }
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (shadow) {
            const __half2 s01 = __floats2half2_rn(pf[0], pf[1]), s23 = __floats2half2_rn(pf[2], pf[3]);
            shadow[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&s01), *reinterpret_cast<const uint32_t*>(&s23));
        }
        if (zero_grad) g[i] = make_uint2(0, 0);
    }
}

// GradScaler.update(): backoff on inf, growth after `growth_interval` clean steps; advances the step count
__global__ void k_scaler_update(ScalerState* st, float growth, float backoff, int growth_interval) {
    if (st->found_inf) {
        st->scale *= backoff;
        st->growth_tracker = 0;
    } else {
        st->step += 1;
        if (++st->growth_tracker >= growth_interval) { st->scale *= growth; st->growth_tracker = 0; }
    }
    st->found_inf = 0;
}

}  // namespace ngp

using namespace ngp;

// state: 8 x 32-bit words {float scale, int growth_tracker, int found_inf, int step, float lr_scale, reserved x3} on the device
extern "C" int ngp_optim_check_finite(const void* grads, int dtype, uint64_t n, void* state, ngp_stream_t stream) {
    if (n == 0) return NGP_OK;
    const uint32_t blocks = (uint32_t)((n + 256 * 8 - 1) / (256 * 8));
    const uint32_t grid = blocks < (uint32_t)sm_count() * 8 ? (blocks ? blocks : 1) : (uint32_t)sm_count() * 8;
    if (dtype == NGP_F16) k_check_finite<__half><<<grid, 256, 0, as_stream(stream)>>>((const __half*)grads, n, (ScalerState*)state);
    else k_check_finite<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)grads, n, (ScalerState*)state);
    return check_launch("optim_check_finite");
}

extern "C" int ngp_optim_adam_step(float* params, float* exp_avg, float* exp_avg_sq, void* grads, int dtype,
                                   void* shadow_f16, uint64_t n, float lr, float beta1, float beta2, float eps,
                                   const void* state, int zero_grad, ngp_stream_t stream) {
    if (n == 0) return NGP_OK;
    const uint32_t blocks = (uint32_t)((n + 256 * 4 - 1) / (256 * 4));
    const uint32_t grid = blocks < (uint32_t)sm_count() * 8 ? (blocks ? blocks : 1) : (uint32_t)sm_count() * 8;
    const bool aligned = ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(grads) & 7) == 0 && (reinterpret_cast<uintptr_t>(shadow_f16) & 7) == 0;
    if (dtype == NGP_F16 && aligned && n % 4 == 0) {
        const size_t n4 = n / 4;
        const uint32_t b4 = (uint32_t)((n4 + 255) / 256);
        const uint32_t g4 = b4 < (uint32_t)sm_count() * 16 ? (b4 ? b4 : 1) : (uint32_t)sm_count() * 16;
        k_adam_step_vec4<<<g4, 256, 0, as_stream(stream)>>>((float4*)params, (float4*)exp_avg, (float4*)exp_avg_sq, (uint2*)grads,
                                                            (uint2*)shadow_f16, n4, lr, beta1, beta2, eps, (const ScalerState*)state, zero_grad);
        return check_launch("optim_adam_step");
    }
    if (dtype == NGP_F16)
        k_adam_step<__half><<<grid, 256, 0, as_stream(stream)>>>(params, exp_avg, exp_avg_sq, (__half*)grads, (__half*)shadow_f16, n, lr,
                                                                 beta1, beta2, eps, (const ScalerState*)state, zero_grad);
    else
        k_adam_step<float><<<grid, 256, 0, as_stream(stream)>>>(params, exp_avg, exp_avg_sq, (float*)grads, (__half*)shadow_f16, n, lr,
                                                                beta1, beta2, eps, (const ScalerState*)state, zero_grad);
    return check_launch("optim_adam_step");
}

extern "C" int ngp_optim_scaler_update(void* state, float growth, float backoff, int growth_interval, ngp_stream_t stream) {
    k_scaler_update<<<1, 1, 0, as_stream(stream)>>>((ScalerState*)state, growth, backoff, growth_interval);
    return check_launch("optim_scaler_update");
}
