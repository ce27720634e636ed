// sh.cu — real spherical-harmonics direction encoding, degree 1..8, for sm_100a.
//
// Replaces shencoder/src/shencoder.cu (kernel_sh :27-355, kernel_sh_backward :358-382).
// The reference spells out 64 closed-form polynomials; this implementation evaluates the same
// basis by recurrence instead:
//     Y_{l,+m} = N_l^m * Q_l^m(z) * Re((x+iy)^m),   Y_{l,-m} = N_l^m * Q_l^m(z) * Im((x+iy)^m)
// with Q_l^m = d^m P_l / dz^m (Legendre derivative, a polynomial in z only — the reference's
// polynomials have exactly this "z-polynomial times xy-harmonic" shape, shencoder.cu:50-121) and
// N_l^m = (-1)^m sqrt((2l+1)/(4pi) (l-m)!/(l+m)!) (* sqrt2 for m>0).  Channel order l*l + l + m.
// The analytic Jacobian uses dQ_l^m/dz = Q_l^{m+1} and d/dx,d/dy of (x+iy)^m = m (x+iy)^{m-1}.
// Results agree with the reference to fp32 rounding (tests: 1e-5 abs), not bit-for-bit.
#include "common.cuh"
#include "sh.cuh"

namespace ngp {

template <int DEG>
__global__ void k_sh_forward(const float* __restrict__ inputs, float* __restrict__ outputs, uint32_t B,
                             uint32_t D, float* __restrict__ dy_dx) {
    const uint32_t b = threadIdx.x + blockIdx.x * blockDim.x;
    if (b >= B) return;
    constexpr int C2 = DEG * DEG;
    const float x = inputs[(size_t)b * D + 0], y = inputs[(size_t)b * D + 1], z = inputs[(size_t)b * D + 2];

    float Q[DEG][DEG + 1];
    legendre_derivs<DEG>(z, Q);
    float re[DEG], im[DEG];      // (x + i y)^m
    re[0] = 1.f; im[0] = 0.f;
#pragma unroll
    for (int m = 1; m < DEG; ++m) {
        re[m] = x * re[m - 1] - y * im[m - 1];
        im[m] = x * im[m - 1] + y * re[m - 1];
    }

    float out[C2];
#pragma unroll
    for (int l = 0; l < DEG; ++l) {
        out[l * l + l] = c_shN[l][0] * Q[l][0];
#pragma unroll
        for (int m = 1; m <= l; ++m) {
            const float nq = c_shN[l][m] * Q[l][m];
            out[l * l + l + m] = nq * re[m];
            out[l * l + l - m] = nq * im[m];
        }
    }
    float* __restrict__ o = outputs + (size_t)b * C2;
    if constexpr (C2 % 4 == 0) {
#pragma unroll
        for (int i = 0; i < C2; i += 4)
            *reinterpret_cast<float4*>(o + i) = make_float4(out[i], out[i + 1], out[i + 2], out[i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < C2; ++i) o[i] = out[i];
    }

    if (dy_dx) {
        float* __restrict__ dx = dy_dx + (size_t)b * D * C2;   // [B, 3, C2]
        float* __restrict__ dy = dx + C2;
        float* __restrict__ dz = dy + C2;
#pragma unroll
        for (int l = 0; l < DEG; ++l) {
            dx[l * l + l] = 0.f;
            dy[l * l + l] = 0.f;
            dz[l * l + l] = c_shN[l][0] * Q[l][1];
#pragma unroll
            for (int m = 1; m <= l; ++m) {
                const float n = c_shN[l][m];
                const float nq = n * Q[l][m];
                const float nq1 = (m + 1 <= l) ? n * Q[l][m + 1] : 0.f;
                const float fm = (float)m;
                // +m channel: Re
                dx[l * l + l + m] = nq * fm * re[m - 1];
                dy[l * l + l + m] = -nq * fm * im[m - 1];
                dz[l * l + l + m] = nq1 * re[m];
                // -m channel: Im
                dx[l * l + l - m] = nq * fm * im[m - 1];
                dy[l * l + l - m] = nq * fm * re[m - 1];
                dz[l * l + l - m] = nq1 * im[m];
            }
        }
    }
}

// dL/dd = sum_ch grad * dy_dx  (shencoder.cu:358-382)
__global__ void k_sh_backward(const float* __restrict__ grad, const float* __restrict__ dy_dx,
                              float* __restrict__ grad_inputs, uint32_t B, uint32_t D, uint32_t C2) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float* g = grad + (size_t)b * C2;
    const float* j = dy_dx + (size_t)b * D * C2 + (size_t)d * C2;
    float r = 0.f;
    for (uint32_t c = 0; c < C2; ++c) r = fmaf(g[c], j[c], r);
    grad_inputs[t] = r;
}

}  // namespace ngp

using namespace ngp;

extern "C" int ngp_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D,
                                     uint32_t degree, float* dy_dx, ngp_stream_t stream) {
    if (B == 0) return NGP_OK;
    if (D != 3) return fail(NGP_EINVAL, "SH encoder only support input dim == 3");
    if (degree < 1 || degree > 8) return fail(NGP_EINVAL, "SH encoder only supports degree in [1, 8]");
    cudaStream_t st = as_stream(stream);
    const uint32_t nb = div_up(B, 128u);
    switch (degree) {
        case 1: k_sh_forward<1><<<nb, 128, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 2: k_sh_forward<2><<<nb, 128, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 3: k_sh_forward<3><<<nb, 128, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 4: k_sh_forward<4><<<nb, 128, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 5: k_sh_forward<5><<<nb, 128, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 6: k_sh_forward<6><<<nb, 128, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 7: k_sh_forward<7><<<nb, 128, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        default: k_sh_forward<8><<<nb, 128, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
    }
    return check_launch("sh_encode_forward");
}

extern "C" int ngp_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D,
                                      uint32_t degree, const float* dy_dx, float* grad_inputs,
                                      ngp_stream_t stream) {
    (void)inputs;
    if (B == 0) return NGP_OK;
    if (!dy_dx || !grad_inputs) return fail(NGP_EINVAL, "sh_encode_backward: dy_dx and grad_inputs are required");
    k_sh_backward<<<div_up(B * D, 256u), 256, 0, as_stream(stream)>>>(grad, dy_dx, grad_inputs, B, D, degree * degree);
    return check_launch("sh_encode_backward");
}
