// exchange.cu — the step's only exchange (one sum-reduction of {d table, d sigma weights, d color weights} per step) done
// over NVLink peer memory and FUSED with the optimizer, one process per GPU.
//
// The reference has no live distributed path; the baseline for this step is "NCCL all-reduce of the fp16 gradient sink, then the
// optimizer pass over all 12.2 M parameters on every rank" (ngp_optim.FusedFieldOptimizer.begin_exchange / apply).  Here the flat
// parameter space is cut into `world` shards (multiples of 8 elements) and every rank does, with plain loads / stores on peer-mapped
// pointers (CUDA IPC; NVSwitch gives every pair full bandwidth):
//
//   barrier 0   every rank's scatter / weight-gradient kernels have finished writing its local fp16 sink
//   k_reduce    rank r reads shard r of ALL sinks (7 x 3.06 MB inbound at world 8), sums in fp32, rounds once to fp16, keeps the
//               result in shard r of its own sink and tests it for inf / nan
//   barrier 1   carries each rank's non-finite flag: afterwards every rank knows "skip this step" (GradScaler semantics) and that
//               its sink has been read by everybody
//   k_adam      Adam on the fp32 masters / moments of shard r only (1/world of the optimizer traffic), the updated fp16 operand
//               copy ("shadow": the hash table the kernels gather from and the MLP weights) is stored to ALL replicas (own + peers,
//               7 x 3.06 MB outbound), the whole local sink is cleared for the next step
//   barrier 2   all replicas' shadows are complete -> the next forward may start
//
// i.e. reduce-scatter + sharded optimizer + all-gather of the fp16 parameters in three kernels and three ~2 us flag barriers,
// instead of all-reduce (2 x 24.5 MB per rank through NCCL's staging) + a full-size optimizer pass per rank.  fp32 masters and
// moments of the other shards are NOT kept current on a rank (FusedFieldOptimizer.gather_master() collects them for checkpoints).
// Barriers are single-warp kernels (st.release.sys / ld.acquire.sys on per-rank slots of each peer's signal pad, monotonic epochs
// kept on the device so that a captured CUDA graph replays correctly); spins are bounded by a wall-clock timeout that raises an
// error word instead of hanging the device.
#include "common.cuh"
#include <string.h>

namespace ngp {

static constexpr uint32_t PAD_WORDS = 4096;        // signal pad: 16 KB of u32
static constexpr uint32_t PAD_SLOT_STRIDE = 64;    // words per barrier slot (up to 64 ranks)
static constexpr uint32_t PAD_EPOCH = 2048;        // local epoch counters [slot]
static constexpr uint32_t PAD_ERROR = 3072;        // != 0: a barrier timed out
static constexpr uint32_t MAX_WORLD = 16;

struct PeerPtrs { void* p[MAX_WORLD]; };

struct ScalerStateX { float scale; int growth_tracker; int found_inf; int step; float lr_scale; int reserved[3]; };

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
// NVLS (multicast) forms, used when the ranks' blocks are also mapped through one NVSwitch multicast object: ONE load returns the sum over
// every replica of the addressed 16 bytes (reduced inside the switch, fp32 accumulation, one rounding to fp16), ONE store lands in every replica
// — inbound reduce traffic and outbound operand traffic drop from (world-1)/world of the table to 1/world of it.
__device__ __forceinline__ uint4 multimem_ld_reduce_h8(const void* mc_addr) {
    uint4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc_addr) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st_16(void* mc_addr, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// one warp: lane p talks to rank p.  flag_in (nullable): a device int whose non-zero-ness is OR-ed over all ranks into *flag_out.
__global__ void __launch_bounds__(32)
k_xchg_barrier(PeerPtrs pads, uint32_t* __restrict__ my_pad, uint32_t rank, uint32_t world, uint32_t slot,
               const int* __restrict__ flag_in, int* __restrict__ flag_out, unsigned long long timeout_ns) {
    const uint32_t lane = threadIdx.x;
    uint32_t epoch = 0;
    if (lane == 0) {
        epoch = my_pad[PAD_EPOCH + slot] + 1u;
        my_pad[PAD_EPOCH + slot] = epoch;
    }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    const uint32_t mine = (flag_in && *flag_in) ? 1u : 0u;
    __threadfence_system();
    uint32_t any = 0;
    if (lane < world) {
        st_release_sys(reinterpret_cast<uint32_t*>(pads.p[lane]) + slot * PAD_SLOT_STRIDE + rank, (epoch << 1) | mine);
        const uint32_t* src = my_pad + slot * PAD_SLOT_STRIDE + lane;
        const unsigned long long t0 = globaltimer_ns();
        uint32_t v = ld_acquire_sys(src);
        while ((v >> 1) < epoch) {
            if (my_pad[PAD_ERROR] != 0u) break;
            if (globaltimer_ns() - t0 > timeout_ns) { atomicExch(my_pad + PAD_ERROR, 1u + slot); break; }
            __nanosleep(64);
            v = ld_acquire_sys(src);
        }
        any = v & 1u;
    }
    any = __any_sync(0xffffffffu, any != 0u) ? 1u : 0u;
    if (lane == 0 && flag_out && any) atomicOr(flag_out, 1);
}

__device__ __forceinline__ bool half2_nonfinite(uint32_t u) {
    return ((u & 0x7c00u) == 0x7c00u) || ((u & 0x7c000000u) == 0x7c000000u);
}

// shard [lo8, lo8 + n8) in units of 8 halves: my_sink[i] = fp16(sum_p fp32(sink_p[i])); non-finite results raise state->found_inf
__global__ void __launch_bounds__(256)
k_xchg_reduce(PeerPtrs sinks, uint4* __restrict__ my_sink, uint32_t rank, uint32_t world, size_t lo8, size_t n8,
              ScalerStateX* __restrict__ st) {
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint4 v[MAX_WORLD];
#pragma unroll
        for (uint32_t p = 0; p < MAX_WORLD; ++p)
            if (p < world) v[p] = reinterpret_cast<const uint4*>(sinks.p[p])[lo8 + i];      // all peers' loads in flight together
#pragma unroll
        for (uint32_t p = 0; p < MAX_WORLD; ++p) {
            if (p < world) {
                const __half2* h = reinterpret_cast<const __half2*>(&v[p]);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(h[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
            }
        }
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __half2 h = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
            ow[k] = *reinterpret_cast<const uint32_t*>(&h);
            bad |= half2_nonfinite(ow[k]);
        }
        my_sink[lo8 + i] = o;
    }
    (void)rank;
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31u) == 0) atomicOr(&st->found_inf, 1);
}

// Adam on the piece [lo8, lo8 + n8) (units of 8 elements) of ONE parameter tensor whose first element has flat index seg_off:
//   p      fp32 master of that tensor (element 0 = flat index seg_off)
//   m, v   flat moment arrays (indexed by flat index)
//   sink   this rank's flat fp16 gradient bucket (holds the reduced gradient in this piece)
//   shadows[q] flat fp16 operand copy of rank q (peer-mapped); every replica receives the updated values
__global__ void __launch_bounds__(256)
k_xchg_adam(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, uint4* __restrict__ sink, PeerPtrs shadows,
            uint32_t world, size_t seg_off, size_t lo8, size_t n8, float lr, float beta1, float beta2, float eps,
            const ScalerStateX* __restrict__ st) {
    const bool skip = st->found_inf != 0;
    const float inv_scale = 1.0f / st->scale;
    const int step = st->step + 1;
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2 = 1.0f - powf(beta2, (float)step);
    const float step_size = lr * st->lr_scale / bc1;
    const float rsqrt_bc2 = rsqrtf(bc2);
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    if (skip) return;             // the sink is cleared by k_xchg_zero; parameters, moments and shadows stay as they are
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f8 = lo8 + i;                          // flat index / 8
        const uint4 gr = sink[f8];
        const __half2* gh = reinterpret_cast<const __half2*>(&gr);
        float4* pp = reinterpret_cast<float4*>(p + (f8 * 8 - seg_off));
        float4* mp = reinterpret_cast<float4*>(m + f8 * 8);
        float4* vp = reinterpret_cast<float4*>(v + f8 * 8);
        uint4 sh;
        uint32_t* shw = reinterpret_cast<uint32_t*>(&sh);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float4 P = pp[q], Mm = mp[q], V = vp[q];
            float* pf = reinterpret_cast<float*>(&P); float* mf = reinterpret_cast<float*>(&Mm); float* vf = reinterpret_cast<float*>(&V);
            const float2 g01 = __half22float2(gh[2 * q]), g23 = __half22float2(gh[2 * q + 1]);
            const float gi[4] = {g01.x * inv_scale, g01.y * inv_scale, g23.x * inv_scale, g23.y * inv_scale};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                mf[k] = fmaf(beta1, mf[k], omb1 * gi[k]);
                vf[k] = fmaf(beta2, vf[k], omb2 * gi[k] * gi[k]);
                pf[k] = pf[k] - step_size * (mf[k] / (sqrtf(vf[k]) * rsqrt_bc2 + eps));
            }
            pp[q] = P; mp[q] = Mm; vp[q] = V;
            const __half2 s01 = __floats2half2_rn(pf[0], pf[1]), s23 = __floats2half2_rn(pf[2], pf[3]);
            shw[2 * q] = *reinterpret_cast<const uint32_t*>(&s01);
            shw[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&s23);
        }
#pragma unroll
        for (uint32_t q = 0; q < MAX_WORLD; ++q)
            if (q < world) reinterpret_cast<uint4*>(shadows.p[q])[f8] = sh;
    }
}

__global__ void __launch_bounds__(256) k_xchg_zero(uint4* __restrict__ sink, size_t n8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x)
        sink[i] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------------------------------------
// Fused form (what FusedFieldOptimizer uses): the three flag barriers ride inside the data kernels, so a step's exchange + optimizer
// is THREE launches instead of eight —
//   k_xchg_reduce_f : [signal "my bucket is complete" | every CTA waits for all ranks' signals] reduce my shard, inf/nan flag
//   k_xchg_adam_f   : [signal found_inf | every CTA waits, ORs the flags] Adam on the pieces of my shard (or skip), shadows stored to
//                     every replica, rest of my bucket cleared, [last CTA to finish signals "my shadow stores are done"]
//   k_xchg_finish   : wait for every rank's "done", GradScaler update
// Epochs live in the local pad (PAD_EPOCH + slot) and are advanced by the LAST CTA of the kernel that used them (ticket counter), so
// all CTAs of a launch read the same value and a captured CUDA graph replays correctly.
static constexpr uint32_t PAD_TICKET = 2560;      // per-kernel completion tickets

__device__ __forceinline__ uint32_t wait_all_ranks(const uint32_t* my_pad_c, uint32_t* my_pad, uint32_t world, uint32_t slot, uint32_t epoch,
                                                   unsigned long long timeout_ns) {
    // executed by one full warp; returns the OR of the ranks' flag bits
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t bit = 0;
    if (lane < world) {
        const uint32_t* src = my_pad_c + slot * PAD_SLOT_STRIDE + lane;
        const unsigned long long t0 = globaltimer_ns();
        uint32_t v = ld_acquire_sys(src);
        while ((v >> 1) < epoch) {
            if (*reinterpret_cast<volatile uint32_t*>(my_pad + PAD_ERROR) != 0u) break;
            if (globaltimer_ns() - t0 > timeout_ns) { atomicExch(my_pad + PAD_ERROR, 1u + slot); break; }
            __nanosleep(32);
            v = ld_acquire_sys(src);
        }
        bit = v & 1u;
    }
    return __any_sync(0xffffffffu, bit != 0u) ? 1u : 0u;
}
__device__ __forceinline__ void signal_all_ranks(const PeerPtrs& pads, uint32_t rank, uint32_t world, uint32_t slot, uint32_t value) {
    const uint32_t lane = threadIdx.x & 31u;
    __threadfence_system();
    if (lane < world) st_release_sys(reinterpret_cast<uint32_t*>(pads.p[lane]) + slot * PAD_SLOT_STRIDE + rank, value);
}
// last CTA of the launch (all others have passed their fence + ticket): returns true in exactly one CTA, for all its threads
__device__ __forceinline__ bool last_cta_done(uint32_t* ticket) {
    __shared__ uint32_t s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const uint32_t t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1u) ? 1u : 0u;
        if (s_last) { *ticket = 0u; __threadfence_system(); }
    }
    __syncthreads();
    return s_last != 0u;
}

__global__ void __launch_bounds__(256)
k_xchg_reduce_f(PeerPtrs pads, PeerPtrs sinks, const uint4* __restrict__ mc_sink, uint32_t rank, uint32_t world, size_t lo8, size_t n8,
                ScalerStateX* __restrict__ st, unsigned long long timeout_ns) {
    uint32_t* my_pad = reinterpret_cast<uint32_t*>(pads.p[rank]);
    const uint32_t epoch = my_pad[PAD_EPOCH + 0] + 1u;
    if (threadIdx.x < 32) {
        if (blockIdx.x == 0) signal_all_ranks(pads, rank, world, 0, epoch << 1);
        wait_all_ranks(my_pad, my_pad, world, 0, epoch, timeout_ns);
    }
    __syncthreads();
    uint4* my_sink = reinterpret_cast<uint4*>(sinks.p[rank]);
    bool bad = false;
    if (mc_sink) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
            const uint4 o = multimem_ld_reduce_h8(mc_sink + lo8 + i);
            bad |= half2_nonfinite(o.x) | half2_nonfinite(o.y) | half2_nonfinite(o.z) | half2_nonfinite(o.w);
            my_sink[lo8 + i] = o;
        }
    } else
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint4 v[MAX_WORLD];
#pragma unroll
        for (uint32_t p = 0; p < MAX_WORLD; ++p)
            if (p < world) v[p] = reinterpret_cast<const uint4*>(sinks.p[p])[lo8 + i];
#pragma unroll
        for (uint32_t p = 0; p < MAX_WORLD; ++p) {
            if (p < world) {
                const __half2* h = reinterpret_cast<const __half2*>(&v[p]);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(h[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
            }
        }
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __half2 h = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
            ow[k] = *reinterpret_cast<const uint32_t*>(&h);
            bad |= half2_nonfinite(ow[k]);
        }
        my_sink[lo8 + i] = o;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31u) == 0) atomicOr(&st->found_inf, 1);
    if (last_cta_done(my_pad + PAD_TICKET + 0) && threadIdx.x == 0) my_pad[PAD_EPOCH + 0] = epoch;
}

struct AdamPieces {            // the pieces of this rank's shard, one per parameter tensor it intersects
    float* p[4];
    unsigned long long seg_off[4], lo8[4], n8[4];
    uint32_t count;
};

__global__ void __launch_bounds__(256)
k_xchg_adam_f(PeerPtrs pads, PeerPtrs shadows, uint4* __restrict__ mc_shadow, uint32_t rank, uint32_t world, AdamPieces pieces, float* __restrict__ m, float* __restrict__ v,
              uint4* __restrict__ sink, size_t shard_lo8, size_t shard_hi8, size_t total8, float lr, float beta1, float beta2, float eps,
              ScalerStateX* __restrict__ st, unsigned long long timeout_ns) {
    __shared__ uint32_t s_skip;
    uint32_t* my_pad = reinterpret_cast<uint32_t*>(pads.p[rank]);
    const uint32_t e1 = my_pad[PAD_EPOCH + 1] + 1u;
    if (threadIdx.x < 32) {
        // the reduce kernel has completed (stream order): this rank's non-finite flag is final
        if (blockIdx.x == 0) signal_all_ranks(pads, rank, world, 1, (e1 << 1) | (st->found_inf != 0 ? 1u : 0u));
        const uint32_t any = wait_all_ranks(my_pad, my_pad, world, 1, e1, timeout_ns);
        if (threadIdx.x == 0) {
            s_skip = any;
            if (any && blockIdx.x == 0) st->found_inf = 1;          // for the scaler update (only this CTA touches it here)
        }
    }
    __syncthreads();
    const bool skip = s_skip != 0u;
    const float inv_scale = 1.0f / st->scale;
    const int step = st->step + 1;
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2 = 1.0f - powf(beta2, (float)step);
    const float step_size = lr * st->lr_scale / bc1;
    const float rsqrt_bc2 = rsqrtf(bc2);
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gstride = (size_t)gridDim.x * blockDim.x;
    for (uint32_t pc = 0; pc < pieces.count; ++pc) {
        float* __restrict__ p = pieces.p[pc];
        const size_t seg_off = pieces.seg_off[pc], lo8 = pieces.lo8[pc], n8 = pieces.n8[pc];
        for (size_t i = gtid; i < n8; i += gstride) {
            const size_t f8 = lo8 + i;
            const uint4 gr = sink[f8];
            sink[f8] = make_uint4(0, 0, 0, 0);               // this element of the bucket is consumed
            if (skip) continue;
            const __half2* gh = reinterpret_cast<const __half2*>(&gr);
            float4* pp = reinterpret_cast<float4*>(p + (f8 * 8 - seg_off));
            float4* mp = reinterpret_cast<float4*>(m + f8 * 8);
            float4* vp = reinterpret_cast<float4*>(v + f8 * 8);
            uint4 sh;
            uint32_t* shw = reinterpret_cast<uint32_t*>(&sh);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float4 P = pp[q], Mm = mp[q], V = vp[q];
                float* pf = reinterpret_cast<float*>(&P); float* mf = reinterpret_cast<float*>(&Mm); float* vf = reinterpret_cast<float*>(&V);
                const float2 g01 = __half22float2(gh[2 * q]), g23 = __half22float2(gh[2 * q + 1]);
                const float gi[4] = {g01.x * inv_scale, g01.y * inv_scale, g23.x * inv_scale, g23.y * inv_scale};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mf[k] = fmaf(beta1, mf[k], omb1 * gi[k]);
                    vf[k] = fmaf(beta2, vf[k], omb2 * gi[k] * gi[k]);
                    pf[k] = pf[k] - step_size * (mf[k] / (sqrtf(vf[k]) * rsqrt_bc2 + eps));
                }
                pp[q] = P; mp[q] = Mm; vp[q] = V;
                const __half2 s01 = __floats2half2_rn(pf[0], pf[1]), s23 = __floats2half2_rn(pf[2], pf[3]);
                shw[2 * q] = *reinterpret_cast<const uint32_t*>(&s01);
                shw[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&s23);
            }
            if (mc_shadow) {
                multimem_st_16(mc_shadow + f8, sh);
            } else {
#pragma unroll
                for (uint32_t q = 0; q < MAX_WORLD; ++q)
                    if (q < world) reinterpret_cast<uint4*>(shadows.p[q])[f8] = sh;
            }
        }
    }
    // the other ranks' shards of my bucket: everybody has read them (barrier 1), clear them for the next step
    for (size_t i = gtid; i < total8; i += gstride)
        if (i < shard_lo8 || i >= shard_hi8) sink[i] = make_uint4(0, 0, 0, 0);
    // my stores into the replicas' shadows are complete once every CTA of this launch is past this point
    if (last_cta_done(my_pad + PAD_TICKET + 1)) {
        const uint32_t e2 = my_pad[PAD_EPOCH + 2] + 1u;
        if (threadIdx.x < 32) signal_all_ranks(pads, rank, world, 2, e2 << 1);
        if (threadIdx.x == 0) { my_pad[PAD_EPOCH + 1] = e1; my_pad[PAD_EPOCH + 2] = e2; }
    }
}

__global__ void __launch_bounds__(32)
k_xchg_finish(PeerPtrs pads, uint32_t rank, uint32_t world, ScalerStateX* __restrict__ st, float growth, float backoff, int growth_interval,
              unsigned long long timeout_ns) {
    uint32_t* my_pad = reinterpret_cast<uint32_t*>(pads.p[rank]);
    const uint32_t e2 = my_pad[PAD_EPOCH + 2];         // advanced by k_xchg_adam_f of this step
    wait_all_ranks(my_pad, my_pad, world, 2, e2, timeout_ns);
    if (threadIdx.x == 0) {                              // GradScaler.update() (same rule as k_scaler_update)
        if (st->found_inf) { st->scale *= backoff; st->growth_tracker = 0; }
        else { st->step += 1; if (++st->growth_tracker >= growth_interval) { st->scale *= growth; st->growth_tracker = 0; } }
        st->found_inf = 0;
    }
}

static uint32_t grid_for(size_t n, uint32_t per_sm) {
    const size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)sm_count() * per_sm;
    return (uint32_t)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

static int fill_ptrs(PeerPtrs& pp, void* const* host_ptrs, uint32_t world, const char* who) {
    if (world == 0 || world > MAX_WORLD) return fail(NGP_EINVAL, "%s: world must be in [1, %u]", who, MAX_WORLD);
    memset(&pp, 0, sizeof(pp));
    for (uint32_t r = 0; r < world; ++r) {
        if (!host_ptrs[r]) return fail(NGP_EINVAL, "%s: peer pointer %u is null", who, r);
        pp.p[r] = host_ptrs[r];
    }
    return NGP_OK;
}

}  // namespace ngp

using namespace ngp;

// ---- peer-visible device memory (CUDA IPC).  The allocation is made with cudaMalloc by this library (the caller's caching allocator
// cannot export sub-allocations), exported as a 64-byte handle that the caller ships to the other processes of the node by whatever
// means it has (torch.distributed.all_gather_object in ngp_dp.PeerExchange), and opened there.  Owner frees, openers close. ----
extern "C" int ngp_peer_alloc(size_t bytes, void** ptr_out, void* handle_out_host) {
    if (!ptr_out || !handle_out_host || bytes == 0) return fail(NGP_EINVAL, "peer_alloc: bad arguments");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(NGP_ECUDA, "peer_alloc: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) { cudaFree(p); return fail(NGP_ECUDA, "peer_alloc: memset failed: %s", cudaGetErrorString(e)); }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(NGP_ECUDA, "peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out_host, &h, sizeof(h));
    *ptr_out = p;
    return NGP_OK;
}
extern "C" int ngp_peer_open(const void* handle_host, void** ptr_out) {
    if (!handle_host || !ptr_out) return fail(NGP_EINVAL, "peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(NGP_ECUDA, "peer_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e)); }
    *ptr_out = p;
    return NGP_OK;
}
extern "C" int ngp_peer_close(void* ptr) {
    if (!ptr) return NGP_OK;
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(NGP_ECUDA, "peer_close: %s", cudaGetErrorString(e)); }
    return NGP_OK;
}
extern "C" int ngp_peer_free(void* ptr) {
    if (!ptr) return NGP_OK;
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(NGP_ECUDA, "peer_free: %s", cudaGetErrorString(e)); }
    return NGP_OK;
}
extern "C" size_t ngp_exchange_pad_bytes(void) { return (size_t)PAD_WORDS * 4; }

// pads_host[world]: every rank's signal pad as mapped in THIS process (own pad at [rank]).  slot in [0, 16).
// flag_in / flag_out: device ints (nullable): *flag_out |= OR over ranks of (*flag_in != 0).
extern "C" int ngp_exchange_barrier(void* const* pads_host, uint32_t rank, uint32_t world, uint32_t slot, const int32_t* flag_in,
                                    int32_t* flag_out, uint32_t timeout_ms, ngp_stream_t stream) {
    PeerPtrs pp;
    int rc = fill_ptrs(pp, pads_host, world, "exchange_barrier");
    if (rc) return rc;
    if (rank >= world || slot >= 16) return fail(NGP_EINVAL, "exchange_barrier: bad rank / slot");
    k_xchg_barrier<<<1, 32, 0, as_stream(stream)>>>(pp, (uint32_t*)pads_host[rank], rank, world, slot, flag_in, flag_out,
                                                   (unsigned long long)(timeout_ms ? timeout_ms : 2000u) * 1000000ull);
    return check_launch("exchange_barrier");
}
// error word of the local pad (host read; synchronises the default stream of the caller's choosing beforehand)
extern "C" int ngp_exchange_error(const void* my_pad, uint32_t* error_out_host) {
    if (!my_pad || !error_out_host) return fail(NGP_EINVAL, "exchange_error: bad arguments");
    cudaError_t e = cudaMemcpy(error_out_host, (const uint32_t*)my_pad + PAD_ERROR, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(NGP_ECUDA, "exchange_error: %s", cudaGetErrorString(e));
    return NGP_OK;
}

// sinks_host[world]: every rank's flat fp16 gradient bucket (n elements, n % 8 == 0) as mapped here.  Reduces the elements
// [lo, lo + count) (multiples of 8) into this rank's bucket and raises state.found_inf on a non-finite sum.
extern "C" int ngp_exchange_reduce(void* const* sinks_host, uint32_t rank, uint32_t world, uint64_t lo, uint64_t count, void* state,
                                   ngp_stream_t stream) {
    PeerPtrs pp;
    int rc = fill_ptrs(pp, sinks_host, world, "exchange_reduce");
    if (rc) return rc;
    if (rank >= world || (lo & 7) || (count & 7) || !state) return fail(NGP_EINVAL, "exchange_reduce: bad arguments");
    if (count == 0) return NGP_OK;
    k_xchg_reduce<<<grid_for(count / 8, 8), 256, 0, as_stream(stream)>>>(pp, (uint4*)sinks_host[rank], rank, world, lo / 8, count / 8,
                                                                        (ScalerStateX*)state);
    return check_launch("exchange_reduce");
}

// One parameter tensor's piece [lo, lo + count) (flat indices, multiples of 8; the tensor's element 0 has flat index seg_off).
extern "C" int ngp_exchange_adam(float* params, float* exp_avg_flat, float* exp_avg_sq_flat, void* my_sink, void* const* shadows_host,
                                 uint32_t world, uint64_t seg_off, uint64_t lo, uint64_t count, float lr, float beta1, float beta2,
                                 float eps, const void* state, ngp_stream_t stream) {
    PeerPtrs pp;
    int rc = fill_ptrs(pp, shadows_host, world, "exchange_adam");
    if (rc) return rc;
    if ((lo & 7) || (count & 7) || (seg_off & 7) || lo < seg_off || !params || !exp_avg_flat || !exp_avg_sq_flat || !my_sink || !state)
        return fail(NGP_EINVAL, "exchange_adam: bad arguments");
    if (count == 0) return NGP_OK;
    k_xchg_adam<<<grid_for(count / 8, 8), 256, 0, as_stream(stream)>>>(params, exp_avg_flat, exp_avg_sq_flat, (uint4*)my_sink, pp, world,
                                                                      seg_off, lo / 8, count / 8, lr, beta1, beta2, eps,
                                                                      (const ScalerStateX*)state);
    return check_launch("exchange_adam");
}

extern "C" int ngp_exchange_zero(void* my_sink, uint64_t n, ngp_stream_t stream) {
    if (!my_sink || (n & 7)) return fail(NGP_EINVAL, "exchange_zero: bad arguments");
    if (n == 0) return NGP_OK;
    k_xchg_zero<<<grid_for(n / 8, 8), 256, 0, as_stream(stream)>>>((uint4*)my_sink, n / 8);
    return check_launch("exchange_zero");
}

// ---- fused form: three launches per step (see the kernels' comment) ----
extern "C" int ngp_exchange_reduce_fused(void* const* pads_host, void* const* sinks_host, const void* mc_sink, uint32_t rank, uint32_t world,
                                         uint64_t lo, uint64_t count, void* state, uint32_t timeout_ms, ngp_stream_t stream) {
    PeerPtrs pp, sp;
    int rc = fill_ptrs(pp, pads_host, world, "exchange_reduce_fused");
    if (rc) return rc;
    rc = fill_ptrs(sp, sinks_host, world, "exchange_reduce_fused");
    if (rc) return rc;
    if (rank >= world || (lo & 7) || (count & 7) || !state) return fail(NGP_EINVAL, "exchange_reduce_fused: bad arguments");
    // every CTA spins on the ranks' flags before it reduces: the grid must be co-resident (<= 8 CTAs of 256 threads per SM)
    k_xchg_reduce_f<<<grid_for(count / 8 + 1, 4), 256, 0, as_stream(stream)>>>(pp, sp, (const uint4*)mc_sink, rank, world, lo / 8, count / 8, (ScalerStateX*)state,
                                                                             (unsigned long long)(timeout_ms ? timeout_ms : 2000u) * 1000000ull);
    return check_launch("exchange_reduce_fused");
}

// params_host / seg_off_host / lo_host / count_host: n_pieces (<= 4) parameter pieces of this rank's shard [shard_lo, shard_hi);
// n_total = size of the flat bucket.  Also clears the bucket and signals completion of the shadow stores.
extern "C" int ngp_exchange_adam_fused(void* const* pads_host, void* const* shadows_host, void* mc_shadow, uint32_t rank, uint32_t world,
                                       float* const* params_host, const uint64_t* seg_off_host, const uint64_t* lo_host,
                                       const uint64_t* count_host, uint32_t n_pieces, float* exp_avg_flat, float* exp_avg_sq_flat,
                                       void* my_sink, uint64_t shard_lo, uint64_t shard_hi, uint64_t n_total, float lr, float beta1,
                                       float beta2, float eps, void* state, uint32_t timeout_ms, ngp_stream_t stream) {
    PeerPtrs pp, sh;
    int rc = fill_ptrs(pp, pads_host, world, "exchange_adam_fused");
    if (rc) return rc;
    rc = fill_ptrs(sh, shadows_host, world, "exchange_adam_fused");
    if (rc) return rc;
    if (rank >= world || n_pieces > 4 || !exp_avg_flat || !exp_avg_sq_flat || !my_sink || !state || ((shard_lo | shard_hi | n_total) & 7))
        return fail(NGP_EINVAL, "exchange_adam_fused: bad arguments");
    AdamPieces pc;
    memset(&pc, 0, sizeof(pc));
    pc.count = n_pieces;
    size_t work8 = 0;
    for (uint32_t i = 0; i < n_pieces; ++i) {
        if (!params_host[i] || ((seg_off_host[i] | lo_host[i] | count_host[i]) & 7) || lo_host[i] < seg_off_host[i])
            return fail(NGP_EINVAL, "exchange_adam_fused: bad piece %u", i);
        pc.p[i] = params_host[i]; pc.seg_off[i] = seg_off_host[i]; pc.lo8[i] = lo_host[i] / 8; pc.n8[i] = count_host[i] / 8;
        work8 += count_host[i] / 8;
    }
    k_xchg_adam_f<<<grid_for(work8 + 1, 4), 256, 0, as_stream(stream)>>>(pp, sh, (uint4*)mc_shadow, rank, world, pc, exp_avg_flat, exp_avg_sq_flat, (uint4*)my_sink,
                                                                        shard_lo / 8, shard_hi / 8, n_total / 8, lr, beta1, beta2, eps,
                                                                        (ScalerStateX*)state,
                                                                        (unsigned long long)(timeout_ms ? timeout_ms : 2000u) * 1000000ull);
    return check_launch("exchange_adam_fused");
}

extern "C" int ngp_exchange_finish(void* const* pads_host, uint32_t rank, uint32_t world, void* state, float growth, float backoff,
                                   int growth_interval, uint32_t timeout_ms, ngp_stream_t stream) {
    PeerPtrs pp;
    int rc = fill_ptrs(pp, pads_host, world, "exchange_finish");
    if (rc) return rc;
    if (rank >= world || !state) return fail(NGP_EINVAL, "exchange_finish: bad arguments");
    k_xchg_finish<<<1, 32, 0, as_stream(stream)>>>(pp, rank, world, (ScalerStateX*)state, growth, backoff, growth_interval,
                                                  (unsigned long long)(timeout_ms ? timeout_ms : 2000u) * 1000000ull);
    return check_launch("exchange_finish");
}
