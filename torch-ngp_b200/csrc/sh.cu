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

namespace ngp {

static constexpr int SH_MAX = 8;

// N_l^m, sign folded in (generated from the closed form above in double precision)
__constant__ float c_shN[SH_MAX][SH_MAX] = {
    {0.28209479177387814f, 0, 0, 0, 0, 0, 0, 0},
    {0.48860251190291992f, -0.48860251190291998f, 0, 0, 0, 0, 0, 0},
    {0.63078313050504009f, -0.36418281019735976f, 0.18209140509867988f, 0, 0, 0, 0, 0},
    {0.7463526651802308f, -0.3046971996429772f, 0.096353714754685155f, -0.039336239328442907f, 0, 0, 0, 0},
    {0.84628437532163447f, -0.26761861742291571f, 0.063078313050504001f, -0.016858388283618388f, 0.0059603403376112026f, 0, 0, 0},
    {0.9356025796273888f, -0.24157154730437169f, 0.045652731285460234f, -0.0093188247511476283f, 0.0021964680580751762f, -0.00069458418713245519f, 0, 0},
    {1.0171072362820548f, -0.22195099524523101f, 0.03509353369580661f, -0.0058489222826344353f, 0.0010678622237644956f, -0.00022766899107568562f, 6.5722376641838803e-05f, 0},
    {1.0925484305920792f, -0.20647224590289676f, 0.028097313806030647f, -0.0039735602250741348f, 0.00059903674311141165f, -9.9839457185235285e-05f, 1.9580128477462541e-05f, -5.233009453691466e-06f},
};

// Q[l][m] for l < DEG, m <= l (+ one extra m column for the z-derivative)
template <int DEG>
__device__ __forceinline__ void legendre_derivs(float z, float Q[DEG][DEG + 1]) {
#pragma unroll
    for (int l = 0; l < DEG; ++l)
#pragma unroll
        for (int m = 0; m <= DEG; ++m) Q[l][m] = 0.f;
    float dfact = 1.f;  // (2m-1)!!
#pragma unroll
    for (int m = 0; m < DEG; ++m) {
        if (m > 0) dfact *= (float)(2 * m - 1);
        Q[m][m] = dfact;
        if (m + 1 < DEG) Q[m + 1][m] = (float)(2 * m + 1) * z * dfact;
#pragma unroll
        for (int l = m + 2; l < DEG; ++l)
            Q[l][m] = ((float)(2 * l - 1) * z * Q[l - 1][m] - (float)(l + m - 1) * Q[l - 2][m]) * (1.0f / (float)(l - m));
    }
}

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
