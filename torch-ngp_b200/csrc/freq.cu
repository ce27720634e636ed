// freq.cu — frequency (positional) encoding, the reference's `freqencoder` extension (freqencoder/src/freqencoder.cu:28-104).
//   outputs[b] = [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), cos(2^1 x), ...]   (D-wide blocks, C = D + 2*deg*D columns)
// Not on the --ff hot path ("for breadth"); HBM-bound streaming kernel: one thread per point, the D inputs are
// read once and each output row is written as contiguous floats (the reference launches one thread per output ELEMENT, re-reading
// the input 1 + 2*deg times).  sin/cos use the same fast intrinsic and phase-shift formulation as the reference
// (__sinf(scalbnf(x, f) + phase)), so results are bit-identical to it on the GPU.
#include "common.cuh"

namespace ngp {

static constexpr uint32_t FREQ_TPB = 128;
static constexpr uint32_t FREQ_MAX_D = 8;

__global__ void __launch_bounds__(FREQ_TPB) k_freq_forward(const float* __restrict__ inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                                           float* __restrict__ outputs) {
    const uint32_t b = blockIdx.x * FREQ_TPB + threadIdx.x;
    if (b >= B) return;
    float x[FREQ_MAX_D];
    for (uint32_t d = 0; d < D; ++d) x[d] = __ldg(inputs + (size_t)b * D + d);
    float* o = outputs + (size_t)b * C;
    for (uint32_t d = 0; d < D; ++d) o[d] = x[d];
    o += D;
    const float half_pi = 3.141592653589793f / 2;       // the reference's (col % 2) * (PI() / 2)
    for (uint32_t f = 0; f < deg; ++f) {
        for (uint32_t d = 0; d < D; ++d) o[d] = __sinf(scalbnf(x[d], (int)f) + 0.0f);
        for (uint32_t d = 0; d < D; ++d) o[D + d] = __sinf(scalbnf(x[d], (int)f) + half_pi);
        o += 2 * D;
    }
}

// d out / d x: sin block -> 2^f cos, cos block -> -2^f sin, both read back from `outputs` (freqencoder.cu:66-104)
__global__ void __launch_bounds__(FREQ_TPB) k_freq_backward(const float* __restrict__ grad, const float* __restrict__ outputs, uint32_t B,
                                                            uint32_t D, uint32_t deg, uint32_t C, float* __restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * FREQ_TPB + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float* g = grad + (size_t)b * C;
    const float* y = outputs + (size_t)b * C;
    float result = g[d];
    g += D; y += D;
    for (uint32_t f = 0; f < deg; ++f) {
        result += scalbnf(1.0f, (int)f) * (g[d] * y[D + d] - g[D + d] * y[d]);
        g += 2 * D; y += 2 * D;
    }
    grad_inputs[t] = result;
}

}  // namespace ngp

using namespace ngp;

extern "C" int ngp_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs,
                                       ngp_stream_t stream) {
    if (D == 0 || D > FREQ_MAX_D) return fail(NGP_EUNSUPPORTED, "freq_encode_forward: input_dim must be in [1, %u]", FREQ_MAX_D);
    if (C != D + 2 * deg * D) return fail(NGP_EINVAL, "freq_encode_forward: output_dim must be D + 2*degree*D");
    if (B == 0) return NGP_OK;
    k_freq_forward<<<div_up(B, FREQ_TPB), FREQ_TPB, 0, as_stream(stream)>>>(inputs, B, D, deg, C, outputs);
    return check_launch("freq_encode_forward");
}

extern "C" int ngp_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                        float* grad_inputs, ngp_stream_t stream) {
    if (D == 0 || D > FREQ_MAX_D) return fail(NGP_EUNSUPPORTED, "freq_encode_backward: input_dim must be in [1, %u]", FREQ_MAX_D);
    if (C != D + 2 * deg * D) return fail(NGP_EINVAL, "freq_encode_backward: output_dim must be D + 2*degree*D");
    if ((uint64_t)B * D > 0xffffffffull) return fail(NGP_EINVAL, "freq_encode_backward: too many points");
    if (B == 0) return NGP_OK;
    k_freq_backward<<<div_up(B * D, FREQ_TPB), FREQ_TPB, 0, as_stream(stream)>>>(grad, outputs, B, D, deg, C, grad_inputs);
    return check_launch("freq_encode_backward");
}
