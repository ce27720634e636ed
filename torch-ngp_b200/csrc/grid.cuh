// grid.cuh — device helpers of the hash-grid encoder shared by grid.cu and the fused encoder->MLP kernel.
#pragma once
#include "common.cuh"

namespace ngp {

static constexpr uint32_t TILE_PTS = 32;   // points per CTA tile == warp width

__device__ __forceinline__ float smoothstep_f(float v) { return v * v * (3.0f - 2.0f * v); }
__device__ __forceinline__ float smoothstep_df(float v) { return 6 * v * (1.0f - v); }

// xor-prime spatial hash (instant-ngp's published primes; gridencoder.cu:50-63)
template <uint32_t D>
__device__ __forceinline__ uint32_t hash_coords(const uint32_t p[D]) {
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u,
                                    2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
#pragma unroll
    for (uint32_t i = 0; i < D; ++i) r ^= p[i] * primes[i];
    return r;
}

// entry index inside one level (gridencoder.cu:66-84); returned WITHOUT the *C+ch.
template <uint32_t D>
__device__ __forceinline__ uint32_t level_index(uint32_t gridtype, bool align_corners,
                                                uint32_t hashmap_size, uint32_t resolution,
                                                const uint32_t p[D]) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        if (stride <= hashmap_size) {
            index += p[d] * stride;
            stride *= align_corners ? resolution : (resolution + 1);
        }
    }
    if (gridtype == 0 && stride > hashmap_size) index = hash_coords<D>(p);
    return index % hashmap_size;
}

__device__ __forceinline__ float level_scale(uint32_t level, float S, uint32_t H) {
    return fmaf(exp2f(level * S), (float)H, -1.0f);
}

// All 2^D corner entry indices of one (point, level) at once.  Same results as level_index() (the reference's
// get_grid_index, gridencoder.cu:66-84) but the level's addressing mode is decided once per warp:
//   dense  : every stride fits ((res+1)^D <= size)  -> base + constant corner offsets, no modulo needed
//   hash^2 : hashed level with a power-of-two table  -> per-dim products once, xor + mask per corner
//   generic: anything else (tiled wrap-around, non-power-of-two hashed sizes) -> reference loop
template <uint32_t D>
__device__ __forceinline__ void corner_indices(uint32_t gridtype, bool align_corners, uint32_t hashmap_size,
                                               uint32_t resolution, const uint32_t pg[D], uint32_t out[1u << D]) {
    const uint32_t r1 = align_corners ? resolution : resolution + 1;
    // does the dense index space fit?  (64-bit so (res+1)^D cannot wrap)
    unsigned long long cells = 1;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) cells = cells * r1 > 0xffffffffull ? 0x100000000ull : cells * r1;
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    if (cells <= hashmap_size) {
        uint32_t stride[D], base = 0, st = 1;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) { stride[d] = st; base += pg[d] * st; st *= r1; }
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << D); ++idx) {
            uint32_t v = base;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) if (idx & (1u << d)) v += stride[d];
            out[idx] = v;
        }
    } else if (gridtype == 0 && (hashmap_size & (hashmap_size - 1)) == 0) {
        // NOTE: the reference hashes only if the stride product overflowed the table, which is exactly cells > size
        uint32_t h0[D], h1[D];
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) { h0[d] = pg[d] * primes[d]; h1[d] = h0[d] + primes[d]; }
        const uint32_t mask = hashmap_size - 1;
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << D); ++idx) {
            uint32_t v = 0;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) v ^= (idx & (1u << d)) ? h1[d] : h0[d];
            out[idx] = v & mask;
        }
    } else {
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << D); ++idx) {
            uint32_t pl[D];
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) pl[d] = pg[d] + ((idx >> d) & 1u);
            out[idx] = level_index<D>(gridtype, align_corners, hashmap_size, resolution, pl);
        }
    }
}

__device__ __forceinline__ float corner_weight_1(bool hi, float p) { return hi ? p : 1 - p; }

// ---- accumulate / load / store helpers per table dtype --------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// two channels at once: fp32 products, one cvt.rn.f16x2.f32, one HADD2.  Bit-identical to the scalar sequence
// half(float(r) + float(half(w*g))): the sum of two fp16 values is exact in fp32 unless the smaller one is below a
// quarter ulp of the larger, in which case both roundings return the larger operand.
__device__ __forceinline__ void acc2(__half2& r, float w, __half2 g) {
    const float2 gf = __half22float2(g);
    r = __hadd2(r, __floats2half2_rn(w * gf.x, w * gf.y));
}
__device__ __forceinline__ void acc(float& r, float w, float g) { r = fmaf(w, g, r); }
__device__ __forceinline__ void acc(__half& r, float w, __half g) {
    const __half p = __float2half_rn(w * __half2float(g));
    r = __float2half_rn(__half2float(r) + __half2float(p));
}

// vector load of the C features of one table entry
template <typename T, uint32_t C>
__device__ __forceinline__ void load_entry(const T* __restrict__ p, T out[C]) {
    if constexpr (sizeof(T) * C == 4) {
        uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
        *reinterpret_cast<uint32_t*>(out) = v;
    } else if constexpr (sizeof(T) * C == 8) {
        uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
        *reinterpret_cast<uint2*>(out) = v;
    } else if constexpr (sizeof(T) * C == 16) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        *reinterpret_cast<uint4*>(out) = v;
    } else if constexpr (sizeof(T) * C == 32) {
        uint4 v0 = __ldg(reinterpret_cast<const uint4*>(p));
        uint4 v1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        reinterpret_cast<uint4*>(out)[0] = v0;
        reinterpret_cast<uint4*>(out)[1] = v1;
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) out[c] = p[c];
    }
}


}  // namespace ngp
