// sh.cuh — spherical-harmonics device helpers shared by sh.cu and the fused SH->color-MLP kernel.
#pragma once
#include "common.cuh"

namespace ngp {

static constexpr int SH_MAX = 8;

// N_l^m, sign folded in (generated from the closed form above in double precision)
static __device__ const float c_shN[SH_MAX][SH_MAX] = {
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


}  // namespace ngp
