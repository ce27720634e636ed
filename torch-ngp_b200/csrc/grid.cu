// grid.cu — multiresolution hash/tiled grid encoding for sm_100a.
//
// Replaces the reference's gridencoder/src/gridencoder.cu (kernel_grid :87-245,
// kernel_grid_backward :248-340, kernel_input_backward :343-369, kernel_grad_tv :506-610).
// Arithmetic contract kept from the reference:
//   * level scale   = fma(exp2f(level*S), H, -1)           (gridencoder.cu:138)
//   * pos           = fma(x, scale, align ? 0 : 0.5)        (:148)  -> floorf -> frac
//   * corner index  = dense sum while stride <= hashmap_size, else xor-prime hash (:50-84)
//   * fp16 tables   : per-corner product rounded to fp16, then fp16 running sum (:164,187)
//   * fp32 tables   : fused multiply-add running sum
// What is different (B200-first):
//   * one CTA owns a tile of 32 points x all L levels (warp = level, lane = point) so the
//     [B, L*C] feature row is assembled in shared memory and leaves the SM as full 128-bit
//     coalesced stores — the reference's [L,B,C] write + permute copy (grid.py:57) is gone,
//   * the backward reads dL/dy straight from [B, L*C] (no permute copy, grid.py:75) and scatters
//     with red.global.add.noftz.f16x2 / red.global.add.v2.f32 (no return value, one op per corner),
//   * everything runs on the caller's stream.
#include "common.cuh"
#include <stdlib.h>
#include "grid.cuh"

namespace ngp {

// ============================== forward ======================================================
// grid: ceil(B / 32) CTAs; block: 32 x NW threads (NW = min(L, 16) warps); warp w walks levels
// w, w+NW, ...; lane = point.  Dynamic smem: 32 * L * C * sizeof(T) (feature tile).
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(512)
k_grid_forward(const float* __restrict__ inputs, const T* __restrict__ table,
               const int* __restrict__ offsets, T* __restrict__ outputs, const uint32_t B,
               const uint32_t L, const float S, const uint32_t H, T* __restrict__ dy_dx,
               const uint32_t gridtype, const bool align_corners, const uint32_t interp,
               const bool level_major) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);            // [32][pitch] (odd word pitch: conflict-free)
    const uint32_t lane = threadIdx.x;
    const uint32_t warp = threadIdx.y;
    const uint32_t nwarp = blockDim.y;
    const uint32_t b0 = blockIdx.x * TILE_PTS;
    const uint32_t b = b0 + lane;
    const bool valid = b < B;
    const uint32_t F = L * C;
    // row pitch of the staging tile: an odd number of 32-bit words so that the 32 lanes of a warp
    // (same level, consecutive points) hit 32 different banks.
    const bool word_rows = (F * sizeof(T)) % 4 == 0;
    const uint32_t row_words = (F * (uint32_t)sizeof(T)) / 4;
    const uint32_t pitch_words = row_words | 1u;
    const uint32_t pitchE = word_rows ? pitch_words * 4 / (uint32_t)sizeof(T) : F;

    float x[D];
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        x[d] = valid ? __ldg(inputs + (size_t)b * D + d) : 0.f;
        if (x[d] < 0 || x[d] > 1) oob = true;
    }

    for (uint32_t level = warp; level < L; level += nwarp) {
        T res[C];
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) res[c] = from_f<T>(0.f);

        if (valid) {
            const uint32_t off = (uint32_t)__ldg(offsets + level);
            const uint32_t hashmap_size = (uint32_t)__ldg(offsets + level + 1) - off;
            const float scale = level_scale(level, S, H);
            const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
            const T* __restrict__ lvl = table + (size_t)off * C;

            float pos[D], pos_deriv[D];
            uint32_t pg[D];
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) {
                pos[d] = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
                const float fl = floorf(pos[d]);
                pg[d] = (uint32_t)fl;
                pos[d] -= (float)pg[d];
                if (interp == 1) {
                    pos_deriv[d] = smoothstep_df(pos[d]);
                    pos[d] = smoothstep_f(pos[d]);
                } else {
                    pos_deriv[d] = 1.0f;
                }
            }

            if (!oob) {
                // issue all 2^D gathers first (memory-level parallelism), then blend in the
                // reference's corner order (idx bit d selects +1 along dim d).
                T val[1 << D][C];
                float wgt[1 << D];
                uint32_t cidx[1 << D];
                corner_indices<D>(gridtype, align_corners, hashmap_size, resolution, pg, cidx);
#pragma unroll
                for (uint32_t idx = 0; idx < (1u << D); ++idx) load_entry<T, C>(lvl + (size_t)cidx[idx] * C, val[idx]);
#pragma unroll
                for (uint32_t idx = 0; idx < (1u << D); ++idx) {
                    float w = 1;     // same multiplication order as the reference: ((1*a0)*a1)*a2
#pragma unroll
                    for (uint32_t d = 0; d < D; ++d) w *= ((idx & (1u << d)) == 0) ? (1 - pos[d]) : pos[d];
                    wgt[idx] = w;
                }
                if constexpr (sizeof(T) == 2 && C % 2 == 0) {
                    __half2* r2 = reinterpret_cast<__half2*>(res);
#pragma unroll
                    for (uint32_t idx = 0; idx < (1u << D); ++idx) {
#pragma unroll
                        for (uint32_t c = 0; c < C / 2; ++c) acc2(r2[c], wgt[idx], reinterpret_cast<const __half2*>(val[idx])[c]);
                    }
                } else {
#pragma unroll
                    for (uint32_t idx = 0; idx < (1u << D); ++idx) {
#pragma unroll
                        for (uint32_t c = 0; c < C; ++c) acc(res[c], wgt[idx], val[idx][c]);
                    }
                }

                if (dy_dx) {
                    T* __restrict__ dst = dy_dx + (size_t)b * D * F + (size_t)level * D * C;  // [B,L,D,C]
#pragma unroll
                    for (uint32_t gd = 0; gd < D; ++gd) {
                        T rg[C];
#pragma unroll
                        for (uint32_t c = 0; c < C; ++c) rg[c] = from_f<T>(0.f);
#pragma unroll
                        for (uint32_t idx = 0; idx < (1u << (D - 1)); ++idx) {
                            float w = scale;
                            uint32_t pl[D];
#pragma unroll
                            for (uint32_t nd = 0; nd < D - 1; ++nd) {
                                const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                                if ((idx & (1u << nd)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                                else                         { w *= pos[d];     pl[d] = pg[d] + 1; }
                            }
                            pl[gd] = pg[gd];
                            const uint32_t il = level_index<D>(gridtype, align_corners, hashmap_size, resolution, pl);
                            pl[gd] = pg[gd] + 1;
                            const uint32_t ir = level_index<D>(gridtype, align_corners, hashmap_size, resolution, pl);
                            T vl[C], vr[C];
                            load_entry<T, C>(lvl + (size_t)il * C, vl);
                            load_entry<T, C>(lvl + (size_t)ir * C, vr);
#pragma unroll
                            for (uint32_t c = 0; c < C; ++c) {
                                // (right - left) is formed in table precision, as the reference does
                                const T diff = from_f<T>(to_f(vr[c]) - to_f(vl[c]));
                                if constexpr (sizeof(T) == 4) {
                                    rg[c] = fmaf(w * to_f(diff), pos_deriv[gd], rg[c]);
                                } else {
                                    const T p = from_f<T>(w * to_f(diff) * pos_deriv[gd]);
                                    rg[c] = from_f<T>(to_f(rg[c]) + to_f(p));
                                }
                            }
                        }
#pragma unroll
                        for (uint32_t c = 0; c < C; ++c) dst[gd * C + c] = rg[c];
                    }
                }
            } else if (dy_dx) {
                T* __restrict__ dst = dy_dx + (size_t)b * D * F + (size_t)level * D * C;
#pragma unroll
                for (uint32_t i = 0; i < D * C; ++i) dst[i] = from_f<T>(0.f);
            }
        }

        if (level_major) {
            if (valid) {
                T* __restrict__ dst = outputs + ((size_t)level * B + b) * C;
#pragma unroll
                for (uint32_t c = 0; c < C; ++c) dst[c] = res[c];
            }
        } else {
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) tile[lane * pitchE + level * C + c] = res[c];
        }
    }

    if (!level_major) {
        __syncthreads();
        // the tile [32][F] is one contiguous span of the [B, F] output
        const uint32_t npts = min(TILE_PTS, B - b0);
        const size_t base = (size_t)b0 * F;
        const uint32_t nelem = npts * F;
        const uint32_t tid = threadIdx.y * 32 + threadIdx.x;
        const uint32_t nthr = blockDim.y * 32;
        if (word_rows) {
            // 32 lanes x 4 B = one full 128-byte line per warp store
            const uint32_t* src = reinterpret_cast<const uint32_t*>(tile);
            uint32_t* dst = reinterpret_cast<uint32_t*>(outputs + base);
            const uint32_t nwords = npts * row_words;
            for (uint32_t i = tid; i < nwords; i += nthr) {
                const uint32_t r = i / row_words, wd = i - r * row_words;
                dst[i] = src[r * pitch_words + wd];
            }
        } else {
            for (uint32_t i = tid; i < nelem; i += nthr) outputs[base + i] = tile[i];
        }
    }
}

// ============================== backward =====================================================
__device__ __forceinline__ void red_add_h2(__half* addr, __half2 v) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(&v);
    asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(addr), "r"(u) : "memory");
}
__device__ __forceinline__ void red_add_f2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
// two adjacent table entries (C == 2) in one L2 reduction: 8-byte aligned f16x4 / 16-byte aligned f32x4
__device__ __forceinline__ void red_add_h4(__half* addr, __half2 lo, __half2 hi) {
    const uint32_t a = *reinterpret_cast<const uint32_t*>(&lo), b = *reinterpret_cast<const uint32_t*>(&hi);
    asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(addr), "r"(a), "r"(b) : "memory");
}
// four adjacent C == 2 half entries (one 16-byte aligned quad) in one reduction: REDG.E.ADD.F16x8
__device__ __forceinline__ void red_add_h8(__half* addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void red_add_f4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f1(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// same tiling as the forward: warp = level, lane = point.  grad is [B, L*C] (or [L,B,C]).
// Samples arrive ordered along rays, so on the coarse levels many consecutive lanes of a warp hit the
// SAME table entry.  Before touching memory, each corner does a segmented warp reduction over runs of
// equal entry index (shuffle scan, fp32) and only the last lane of a run issues the reduction op:
// the number of L2 atomics on the coarse levels drops by the run length, and the addends are summed
// in fp32 before the single fp16 rounding (more accurate than one fp16 atomic per sample).
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// everything about a level that does not depend on the point: computed once per warp when the warp keeps its level for the whole
// kernel (L <= warps per CTA — the persistent CTA then never repeats the exp2 / ceil / addressing-mode decision per tile)
struct BwdLevel {
    uint32_t off, size, res, mode;      // mode 0: dense strides, 1: xor-prime hash with a power-of-two table, 2: reference loop
    float scale;
};
template <uint32_t D>
__device__ __forceinline__ BwdLevel make_bwd_level(const int* __restrict__ offsets, uint32_t level, float S, uint32_t H,
                                                   uint32_t gridtype, bool align_corners) {
    BwdLevel P;
    P.off = (uint32_t)__ldg(offsets + level);
    P.size = (uint32_t)__ldg(offsets + level + 1) - P.off;
    P.scale = level_scale(level, S, H);
    P.res = (uint32_t)ceilf(P.scale) + 1;
    const uint32_t r1 = align_corners ? P.res : P.res + 1;
    unsigned long long cells = 1;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) cells = cells * r1 > 0xffffffffull ? 0x100000000ull : cells * r1;
    P.mode = cells <= P.size ? 0u : ((gridtype == 0 && (P.size & (P.size - 1)) == 0) ? 1u : 2u);
    return P;
}
// corner_indices() with the addressing mode already decided (same results)
template <uint32_t D>
__device__ __forceinline__ void corner_indices_mode(const BwdLevel& P, uint32_t gridtype, bool align_corners, const uint32_t pg[D],
                                                    uint32_t out[1u << D]) {
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    if (P.mode == 0u) {
        const uint32_t r1 = align_corners ? P.res : P.res + 1;
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
    } else if (P.mode == 1u) {
        uint32_t h0[D], h1[D];
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) { h0[d] = pg[d] * primes[d]; h1[d] = h0[d] + primes[d]; }
        const uint32_t mask = P.size - 1;
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
            out[idx] = level_index<D>(gridtype, align_corners, P.size, P.res, pl);
        }
    }
}

// Persistent CTAs (grid = min(tiles, 3 x SMs)) walk the 32-point tiles round-robin.  With `staged` the tile's inputs — the dL/dy rows
// [32, L*C] and the coordinates [32, D] — are copied global -> shared with cp.async ONE TILE AHEAD (double buffer), fully coalesced,
// and the 16 level-warps read them from shared memory (odd row pitch: conflict-free).  The unstaged form (level-major gradients,
// entries that are not whole 32-bit words) loads them per warp from global: every warp re-reads the same 32 coordinate triples and
// touches 32 sectors for its 32 x 4-byte gradient entries (ncu r2 of that form: L1/LSU 69 % busy, 29 % of the stall samples on the
// CTA's first loads, issue slots 74 % busy with ~580 instructions per warp-level, ~150 of them per-level setup that is now hoisted).
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(512, (D <= 3 && C <= 2) ? 3 : 1)     // NeRF / SDF shapes: 40 registers, 3 CTAs per SM
k_grid_backward(const T* __restrict__ grad, const float* __restrict__ inputs,
                const int* __restrict__ offsets, T* __restrict__ grad_table, const uint32_t B,
                const uint32_t L, const float S, const uint32_t H, const uint32_t gridtype,
                const bool align_corners, const uint32_t interp, const bool level_major, const bool staged, const uint32_t merge_min) {
    extern __shared__ __align__(16) uint32_t stage_sm[];
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x;
    const uint32_t warp = threadIdx.y;
    const uint32_t nwarp = blockDim.y;
    const uint32_t tid = warp * 32 + lane, nthr = nwarp * 32;
    const uint32_t F = L * C;
    const uint32_t ntiles = div_up(B, TILE_PTS);
    constexpr uint32_t EW = (C * (uint32_t)sizeof(T)) / 4;            // 32-bit words per (level, point) gradient entry (staged form)
    // A CTA iteration handles a GROUP of two consecutive 32-point tiles: warp w works on level w of the first tile and on level
    // L-1-w of the second, so that every warp gets one cheap and one expensive level per iteration (coarse levels pay for deep run
    // scans, fine levels for many reductions: with one tile per iteration a quarter of the stall samples sat at the CTA barrier,
    // ncu r2b) and the staging barrier is crossed once per 64 points.
    constexpr uint32_t GROUP = 2;
    const uint32_t ngroups = div_up(ntiles, GROUP);
    const uint32_t row_words = (F * (uint32_t)sizeof(T)) / 4;
    const uint32_t pitch = row_words | 1u;
    const uint32_t tile_words = TILE_PTS * pitch + TILE_PTS * D;
    const uint32_t buf_words = GROUP * tile_words;
    auto issue = [&](uint32_t group, uint32_t buf) {
        if (group < ngroups) {
            const uint32_t b0 = group * GROUP * TILE_PTS;
            const uint32_t npts = min(GROUP * TILE_PTS, B - b0);
            const uint32_t* gsrc = reinterpret_cast<const uint32_t*>(grad + (size_t)b0 * F);
            for (uint32_t i = tid; i < npts * row_words; i += nthr) {
                const uint32_t r = i / row_words, w = i - r * row_words;
                uint32_t* sg = stage_sm + buf * buf_words + (r / TILE_PTS) * tile_words;
                cp_async_4(sg + (r % TILE_PTS) * pitch + w, gsrc + i);
            }
            const float* xsrc = inputs + (size_t)b0 * D;
            for (uint32_t i = tid; i < npts * D; i += nthr) {
                const uint32_t r = i / D;
                float* sx = reinterpret_cast<float*>(stage_sm + buf * buf_words + (r / TILE_PTS) * tile_words + TILE_PTS * pitch);
                cp_async_4(sx + (i - (r / TILE_PTS) * TILE_PTS * D), xsrc + i);
            }
        }
        cp_async_commit_group();          // one group per call, also when empty: the wait distance below stays fixed
    };
    if (staged) issue(blockIdx.x, 0);
    // per-level constants: computed once per CTA, read back from shared memory per (tile, level)
    constexpr uint32_t MAX_TAB = 64;
    __shared__ BwdLevel s_levels[MAX_TAB];
    const bool tabled = L <= MAX_TAB;
    if (tabled && tid < L) s_levels[tid] = make_bwd_level<D>(offsets, tid, S, H, gridtype, align_corners);
    __syncthreads();
    const bool mirror = L == nwarp;        // second tile of a group: warp w takes level L-1-w

    uint32_t it = 0;
    for (uint32_t group = blockIdx.x; group < ngroups; group += gridDim.x, ++it) {
        if (staged) {
            issue(group + gridDim.x, (it + 1u) & 1u);
            cp_async_wait_group<1>();
            __syncthreads();
        }
#pragma unroll 1
      for (uint32_t half = 0; half < GROUP; ++half) {
        const uint32_t tile = group * GROUP + half;
        const uint32_t b = tile * TILE_PTS + lane;
        bool active = b < B;
        const uint32_t* sg = stage_sm + (it & 1u) * buf_words + half * tile_words;
        const float* sx = reinterpret_cast<const float*>(sg + TILE_PTS * pitch);
        const bool flip = mirror && half == 1u;

        float x[D];
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            x[d] = active ? (staged ? sx[lane * D + d] : __ldg(inputs + (size_t)b * D + d)) : 0.5f;
            if (x[d] < 0 || x[d] > 1) active = false;   // grad_table starts at zero (gridencoder.cu:284-289)
        }
        // every warp of the CTA looks at the same 32 points: the skip is CTA-uniform, the barriers stay matched
        const bool any_active = __ballot_sync(FULL, active) != 0;
        if (!active) {
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) x[d] = 0.5f;
        }

        for (uint32_t lv = warp; any_active && lv < L; lv += nwarp) {
            const uint32_t level = flip ? L - 1u - lv : lv;
            const BwdLevel P = tabled ? s_levels[level] : make_bwd_level<D>(offsets, level, S, H, gridtype, align_corners);
            const uint32_t off = P.off;
            const float scale = P.scale;
            T* __restrict__ lvl = grad_table + (size_t)off * C;

            T g[C];
            if (active) {
                if (staged) {
                    if constexpr (EW > 0) {
                        uint32_t* gw = reinterpret_cast<uint32_t*>(g);
#pragma unroll
                        for (uint32_t k = 0; k < EW; ++k) gw[k] = sg[lane * pitch + level * EW + k];
                    }
                } else {
                    const T* __restrict__ gsrc = level_major ? grad + ((size_t)level * B + b) * C
                                                             : grad + (size_t)b * F + (size_t)level * C;
                    load_entry<T, C>(gsrc, g);
                }
            } else {
#pragma unroll
                for (uint32_t c = 0; c < C; ++c) g[c] = from_f<T>(0.f);
            }

            float pos[D];
            uint32_t pg[D];
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) {
                pos[d] = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
                pg[d] = (uint32_t)floorf(pos[d]);
                pos[d] -= (float)pg[d];
                if (interp == 1) pos[d] = smoothstep_f(pos[d]);
            }

            uint32_t cidx[1 << D];
            corner_indices_mode<D>(P, gridtype, align_corners, pg, cidx);

            // ---- run structure of this level, shared by all 2^D corners: consecutive lanes in the SAME CELL have
            // identical corner indices.  (Inactive lanes are isolated so that a run always ends on an active lane.)
            // NOTE: no short-circuit '&&' here — every lane must execute every shuffle
            bool same = active;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) {
                const uint32_t prev_pg = __shfl_up_sync(FULL, pg[d], 1);
                same = same & (prev_pg == pg[d]);
            }
            const int prev_active = __shfl_up_sync(FULL, (int)active, 1);
            same = same & (prev_active != 0);
            const bool head = (lane == 0) || !same;
            const uint32_t heads = __ballot_sync(FULL, head);
            const uint32_t my_head = 31u - __clz(heads & (FULL >> (31u - lane)));
            // Merging costs issue slots (4 shuffles + 4 predicated adds per corner pair and scan step) and saves reduction lane-ops
            // (LSU / L2).  The kernel is issue-bound (ncu r2: 74 % issue-slot utilisation, reductions at 54 % of the measured L2
            // reduction rate), so runs are merged only where enough lanes fold away to pay for the scan (warp-uniform decision).
            const bool do_merge = (32u - __popc(heads)) >= merge_min;
            const uint32_t maxrun = do_merge ? __reduce_max_sync(FULL, lane - my_head) + 1u : 1u;      // redux.sync: longest run in the warp
            const bool issue_red = active && (!do_merge || (lane == 31u) || ((heads >> (lane + 1u)) & 1u));

            if constexpr (C == 2) {
                // corners 2j and 2j+1 differ only in x.  When their entries are an aligned adjacent pair (dense level with an
                // even base index; hashed level with an even x) both are updated by ONE vector reduction (f16x4 / f32x4):
                // 25 % fewer L2 reduction ops on average.  Vector reductions address relative to the level base: the level's first
                // entry must itself be suitably aligned (always true for tables built by GridEncoder: offsets are multiples of 8
                // entries, grid.py:124).  (A 16-byte "quad" form covering x / x+1 inside one aligned group of 4 entries was measured
                // slower — 1.73 -> 1.80 ms, the selection logic costs more issue slots than the 12 % fewer reductions save — and removed.)
                const bool pair_lvl = (off & 1u) == 0u;
#pragma unroll
                for (uint32_t j = 0; j < (1u << (D - 1)); ++j) {
                    // same multiplication order as the reference: ((1 * a0) * a1) * a2
                    float w0 = 1 - pos[0], w1 = pos[0];
#pragma unroll
                    for (uint32_t d = 1; d < D; ++d) {
                        const float f = (((2 * j) & (1u << d)) == 0) ? (1 - pos[d]) : pos[d];
                        w0 *= f; w1 *= f;
                    }
                    float v0[2] = {w0 * to_f(g[0]), w0 * to_f(g[1])}, v1[2] = {w1 * to_f(g[0]), w1 * to_f(g[1])};
                    for (uint32_t o = 1; o < maxrun; o <<= 1) {
#pragma unroll
                        for (uint32_t c = 0; c < 2; ++c) {
                            const float t0 = __shfl_up_sync(FULL, v0[c], o), t1 = __shfl_up_sync(FULL, v1[c], o);
                            if (lane >= my_head + o) { v0[c] += t0; v1[c] += t1; }
                        }
                    }
                    if (issue_red) {
                        const uint32_t i0 = cidx[2 * j], i1 = cidx[2 * j + 1];
                        const bool pair = ((i0 ^ i1) == 1u) && pair_lvl;
                        if constexpr (sizeof(T) == 2) {
                            const __half2 h0 = __floats2half2_rn(v0[0], v0[1]), h1 = __floats2half2_rn(v1[0], v1[1]);
                            if (pair) {
                                red_add_h4(reinterpret_cast<__half*>(lvl + (size_t)(i0 & ~1u) * 2), (i0 < i1) ? h0 : h1, (i0 < i1) ? h1 : h0);
                            } else {
                                red_add_h2(reinterpret_cast<__half*>(lvl + (size_t)i0 * 2), h0);
                                red_add_h2(reinterpret_cast<__half*>(lvl + (size_t)i1 * 2), h1);
                            }
                        } else {
                            if (pair) {
                                float* dst = reinterpret_cast<float*>(lvl + (size_t)(i0 & ~1u) * 2);
                                if (i0 < i1) red_add_f4(dst, v0[0], v0[1], v1[0], v1[1]);
                                else red_add_f4(dst, v1[0], v1[1], v0[0], v0[1]);
                            } else {
                                red_add_f2(reinterpret_cast<float*>(lvl + (size_t)i0 * 2), v0[0], v0[1]);
                                red_add_f2(reinterpret_cast<float*>(lvl + (size_t)i1 * 2), v1[0], v1[1]);
                            }
                        }
                    }
                }
            } else {
#pragma unroll
                for (uint32_t idx = 0; idx < (1u << D); ++idx) {
                    float w = 1;
#pragma unroll
                    for (uint32_t d = 0; d < D; ++d) w *= ((idx & (1u << d)) == 0) ? (1 - pos[d]) : pos[d];
                    const uint32_t index = cidx[idx];

                    // addends in fp32; segmented inclusive scan over the run, only as deep as the longest run
                    float v[C];
#pragma unroll
                    for (uint32_t c = 0; c < C; ++c) v[c] = w * to_f(g[c]);
                    for (uint32_t o = 1; o < maxrun; o <<= 1) {
#pragma unroll
                        for (uint32_t c = 0; c < C; ++c) {
                            const float t = __shfl_up_sync(FULL, v[c], o);
                            if (lane >= my_head + o) v[c] += t;
                        }
                    }
                    if (issue_red) {
                        T* dst = lvl + (size_t)index * C;
                        if constexpr (sizeof(T) == 2 && (C % 2 == 0)) {
#pragma unroll
                            for (uint32_t c = 0; c < C; c += 2) {
                                __half2 hv;
                                hv.x = __float2half_rn(v[c]);
                                hv.y = __float2half_rn(v[c + 1]);
                                red_add_h2(reinterpret_cast<__half*>(dst + c), hv);
                            }
                        } else if constexpr (sizeof(T) == 2) {
                            // C == 1 half: scalar f16 reduction (the reference's path for this case is a stub)
                            atomicAdd(reinterpret_cast<__half*>(dst), __float2half_rn(v[0]));
                        } else if constexpr (C % 2 == 0) {
#pragma unroll
                            for (uint32_t c = 0; c < C; c += 2) red_add_f2(reinterpret_cast<float*>(dst + c), v[c], v[c + 1]);
                        } else {
                            red_add_f1(reinterpret_cast<float*>(dst), v[0]);
                        }
                    }
                }
            }
        }
      }
        if (staged) __syncthreads();       // everybody is done with this buffer before the next iteration's copy lands in it
    }
}

// dL/dx = sum_{l,c} dL/dy * dy_dx   (gridencoder.cu:343-369); one thread per (point, dim)
template <typename T, uint32_t D, uint32_t C>
__global__ void k_grid_input_backward(const T* __restrict__ grad, const T* __restrict__ dy_dx,
                                      T* __restrict__ grad_inputs, uint32_t B, uint32_t L,
                                      const bool level_major) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const uint32_t F = L * C;
    const T* __restrict__ dd = dy_dx + (size_t)b * L * D * C;
    T result = from_f<T>(0.f);
    for (uint32_t l = 0; l < L; ++l) {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) {
            const T gv = level_major ? grad[((size_t)l * B + b) * C + c] : grad[(size_t)b * F + l * C + c];
            const T dv = dd[l * D * C + d * C + c];
            if constexpr (sizeof(T) == 4) {
                result = fmaf(to_f(gv), to_f(dv), result);
            } else {
                const T p = from_f<T>(to_f(gv) * to_f(dv));
                result = from_f<T>(to_f(result) + to_f(p));
            }
        }
    }
    grad_inputs[t] = result;
}

// total-variation gradient (gridencoder.cu:506-610); cold path, thread per (point, level)
template <typename T, uint32_t D, uint32_t C>
__global__ void k_grid_grad_tv(const T* __restrict__ inputs, const T* __restrict__ table,
                               T* __restrict__ grad, const int* __restrict__ offsets,
                               const float weight, const uint32_t B, const uint32_t L, const float S,
                               const uint32_t H, const uint32_t gridtype, const bool align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const uint32_t off = (uint32_t)offsets[level];
    const T* __restrict__ lvl = table + (size_t)off * C;
    T* __restrict__ glvl = grad + (size_t)off * C;

    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        x[d] = to_f(inputs[(size_t)b * D + d]);
        if (x[d] < 0 || x[d] > 1) return;
    }
    const uint32_t hashmap_size = (uint32_t)offsets[level + 1] - off;
    const float scale = level_scale(level, S, H);
    const uint32_t resolution = (uint32_t)ceilf(scale) + 1;

    uint32_t pg[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) pg[d] = (uint32_t)floorf(fmaf(x[d], scale, align_corners ? 0.0f : 0.5f));

    float results[C], idelta[C];
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) { results[c] = 0.f; idelta[c] = 0.f; }
    const uint32_t index = level_index<D>(gridtype, align_corners, hashmap_size, resolution, pg);
    const float w = weight / (2 * D);
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        const uint32_t cur = pg[d];
        if (cur < resolution) {
            pg[d] = cur + 1;
            const uint32_t ir = level_index<D>(gridtype, align_corners, hashmap_size, resolution, pg);
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) {
                const float gv = to_f(lvl[(size_t)index * C + c]) - to_f(lvl[(size_t)ir * C + c]);
                results[c] += gv; idelta[c] += gv * gv;
            }
        }
        if (cur > 0) {
            pg[d] = cur - 1;
            const uint32_t il = level_index<D>(gridtype, align_corners, hashmap_size, resolution, pg);
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) {
                const float gv = to_f(lvl[(size_t)index * C + c]) - to_f(lvl[(size_t)il * C + c]);
                results[c] += gv; idelta[c] += gv * gv;
            }
        }
        pg[d] = cur;
    }
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) {
        const float v = w * results[c] * rsqrtf(idelta[c] + 1e-9f);
        if constexpr (sizeof(T) == 4) atomicAdd(reinterpret_cast<float*>(glvl + (size_t)index * C + c), v);
        else atomicAdd(reinterpret_cast<__half*>(glvl + (size_t)index * C + c), __float2half_rn(v));
    }
}

// per-level scale table as computed on the device (test hook: lets the CPU oracle use the exact
// ex2.approx-based values the kernels use; see oracle/ngp_oracle.c grid_level_scale()).
__global__ void k_level_scales(float* out, uint32_t L, float S, uint32_t H) {
    const uint32_t l = threadIdx.x;
    if (l < L) out[l] = level_scale(l, S, H);
}

// ---- L2 reduction-rate probe (measurement hook, not part of the reference ABI) -----------------------------------------------
// What bounds k_grid_backward is not HBM: the fp16 gradient table (24.5 MB) lives in L2 and every corner update is an L2 reduction
// op.  This kernel measures the device's sustained rate for exactly that access pattern — `ops_per_thread` reductions per thread at
// pseudo-random, suitably aligned entries of a table of `entries` f16x2 values — for the three op widths the scatter uses
// (mode 0: 4-byte f16x2, 1: 8-byte v2.f16x2, 2: 16-byte v4.f16x2).  bench.py reports the scatter against this measured ceiling.
__global__ void k_red_probe(__half* __restrict__ table, uint32_t entries, uint32_t ops_per_thread, int mode) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const __half2 one = __floats2half2_rn(1.0f, 1.0f);
    const uint32_t u = *reinterpret_cast<const uint32_t*>(&one);
    for (uint32_t i = 0; i < ops_per_thread; ++i) {
        x = x * 1664525u + 1013904223u;                       // LCG: a different random entry per op and per lane
        const uint32_t e = (x >> 8) % entries;
        if (mode == 0) red_add_h2(table + (size_t)e * 2, one);
        else if (mode == 1) red_add_h4(table + (size_t)(e & ~1u) * 2, one, one);
        else red_add_h8(table + (size_t)(e & ~3u) * 2, u, u, u, u);
    }
}

// ---- dispatch -------------------------------------------------------------------------------
template <typename T, uint32_t D, uint32_t C>
static int launch_fwd(const float* inputs, const void* emb, const int* offsets, void* out, uint32_t B,
                      uint32_t L, float S, uint32_t H, void* dy_dx, uint32_t gridtype, bool ac,
                      uint32_t interp, bool level_major, cudaStream_t st) {
    const uint32_t nw = L < 16 ? L : 16;
    dim3 block(32, nw);
    dim3 grid(div_up(B, TILE_PTS));
    size_t smem = level_major ? 0 : (size_t)TILE_PTS * ((((size_t)L * C * sizeof(T)) / 4 | 1) * 4 + 4);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_grid_forward<T, D, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(NGP_EINVAL, "grid_encode_forward: L*C too large for shared memory");
    }
    k_grid_forward<T, D, C><<<grid, block, smem, st>>>(inputs, (const T*)emb, offsets, (T*)out, B, L, S, H,
                                                      (T*)dy_dx, gridtype, ac, interp, level_major);
    return check_launch("grid_encode_forward");
}

template <typename T, uint32_t D, uint32_t C>
static int launch_bwd(const void* grad, const float* inputs, const int* offsets, void* gemb, uint32_t B,
                      uint32_t L, float S, uint32_t H, const void* dy_dx, void* ginp, uint32_t gridtype,
                      bool ac, uint32_t interp, bool level_major, cudaStream_t st) {
    const uint32_t nw = L < 16 ? L : 16;
    dim3 block(32, nw);
    const uint32_t ngroups = div_up(div_up(B, TILE_PTS), 2u);              // a CTA iteration handles two 32-point tiles
    const uint32_t cap = (uint32_t)sm_count() * (nw > 8 ? 3u : 6u);      // resident CTAs per SM at 40 registers / thread
    dim3 grid(ngroups < cap ? ngroups : cap);
    // staged (cp.async double-buffered) inputs: point-major gradients whose (level, point) entries are whole 32-bit words
    const uint32_t row_words = (uint32_t)(((size_t)L * C * sizeof(T)) / 4);
    const size_t smem = 2 * 2 * (size_t)(TILE_PTS * (row_words | 1u) + TILE_PTS * D) * 4;      // 2 buffers x 2 tiles
    static const bool stage_env = [] { const char* e = getenv("NGP_GRID_BWD_STAGED"); return e && e[0] == '1'; }();
    // runs of equal cells are merged by a warp scan only when at least this many of the 32 lanes would fold away (see the kernel)
    static const uint32_t merge_min = [] { const char* e = getenv("NGP_GRID_MERGE_MIN"); return e ? (uint32_t)atoi(e) : 0u; }();
    const bool staged = stage_env && !level_major && (C * sizeof(T)) % 4 == 0 && smem <= 40 * 1024;
    k_grid_backward<T, D, C><<<grid, block, staged ? smem : 0, st>>>((const T*)grad, inputs, offsets, (T*)gemb, B, L, S, H,
                                                                      gridtype, ac, interp, level_major, staged, merge_min);
    int rc = check_launch("grid_encode_backward");
    if (rc) return rc;
    if (dy_dx && ginp) {
        k_grid_input_backward<T, D, C><<<div_up(B * D, 256u), 256, 0, st>>>((const T*)grad, (const T*)dy_dx,
                                                                            (T*)ginp, B, L, level_major);
        rc = check_launch("grid_encode_backward(input)");
    }
    return rc;
}

template <typename T, uint32_t D, uint32_t C>
static int launch_tv(const void* inputs, const void* emb, void* grad, const int* offsets, float weight,
                     uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool ac, cudaStream_t st) {
    dim3 grid(div_up(B, 512u), L);
    k_grid_grad_tv<T, D, C><<<grid, 512, 0, st>>>((const T*)inputs, (const T*)emb, (T*)grad, offsets, weight,
                                                 B, L, S, H, gridtype, ac);
    return check_launch("grad_total_variation");
}

#define NGP_DISPATCH_DC(FN, T, ...)                                                             \
    switch (D * 16 + C) {                                                                       \
        case 2 * 16 + 1: return FN<T, 2, 1>(__VA_ARGS__);                                       \
        case 2 * 16 + 2: return FN<T, 2, 2>(__VA_ARGS__);                                       \
        case 2 * 16 + 4: return FN<T, 2, 4>(__VA_ARGS__);                                       \
        case 2 * 16 + 8: return FN<T, 2, 8>(__VA_ARGS__);                                       \
        case 3 * 16 + 1: return FN<T, 3, 1>(__VA_ARGS__);                                       \
        case 3 * 16 + 2: return FN<T, 3, 2>(__VA_ARGS__);                                       \
        case 3 * 16 + 4: return FN<T, 3, 4>(__VA_ARGS__);                                       \
        case 3 * 16 + 8: return FN<T, 3, 8>(__VA_ARGS__);                                       \
        case 4 * 16 + 1: return FN<T, 4, 1>(__VA_ARGS__);                                       \
        case 4 * 16 + 2: return FN<T, 4, 2>(__VA_ARGS__);                                       \
        case 4 * 16 + 4: return FN<T, 4, 4>(__VA_ARGS__);                                       \
        case 4 * 16 + 8: return FN<T, 4, 8>(__VA_ARGS__);                                       \
        case 5 * 16 + 1: return FN<T, 5, 1>(__VA_ARGS__);                                       \
        case 5 * 16 + 2: return FN<T, 5, 2>(__VA_ARGS__);                                       \
        case 5 * 16 + 4: return FN<T, 5, 4>(__VA_ARGS__);                                       \
        case 5 * 16 + 8: return FN<T, 5, 8>(__VA_ARGS__);                                       \
        default: return fail(NGP_EINVAL, "GridEncoding: D must be 2..5 and C must be 1, 2, 4, or 8 (got D=%u C=%u)", D, C); \
    }

}  // namespace ngp

using namespace ngp;

extern "C" int ngp_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets,
                                       void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                       uint32_t H, void* dy_dx, uint32_t gridtype, int align_corners,
                                       uint32_t interp, int dtype, int level_major, ngp_stream_t stream) {
    if (B == 0) return NGP_OK;
    if (!inputs || !embeddings || !offsets || !outputs) return fail(NGP_EINVAL, "grid_encode_forward: null pointer");
    cudaStream_t st = as_stream(stream);
    if (dtype == NGP_F16) { NGP_DISPATCH_DC(launch_fwd, __half, inputs, embeddings, offsets, outputs, B, L, S, H, dy_dx, gridtype, align_corners != 0, interp, level_major != 0, st) }
    if (dtype == NGP_F32) { NGP_DISPATCH_DC(launch_fwd, float, inputs, embeddings, offsets, outputs, B, L, S, H, dy_dx, gridtype, align_corners != 0, interp, level_major != 0, st) }
    return fail(NGP_EINVAL, "grid_encode_forward: dtype must be NGP_F32 or NGP_F16");
}

extern "C" int ngp_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                                        const int32_t* offsets, void* grad_embeddings, uint32_t B, uint32_t D,
                                        uint32_t C, uint32_t L, float S, uint32_t H, const void* dy_dx,
                                        void* grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp,
                                        int dtype, int level_major, ngp_stream_t stream) {
    (void)embeddings;
    if (B == 0) return NGP_OK;
    if (!grad || !inputs || !offsets || !grad_embeddings) return fail(NGP_EINVAL, "grid_encode_backward: null pointer");
    cudaStream_t st = as_stream(stream);
    if (dtype == NGP_F16) { NGP_DISPATCH_DC(launch_bwd, __half, grad, inputs, offsets, grad_embeddings, B, L, S, H, dy_dx, grad_inputs, gridtype, align_corners != 0, interp, level_major != 0, st) }
    if (dtype == NGP_F32) { NGP_DISPATCH_DC(launch_bwd, float, grad, inputs, offsets, grad_embeddings, B, L, S, H, dy_dx, grad_inputs, gridtype, align_corners != 0, interp, level_major != 0, st) }
    return fail(NGP_EINVAL, "grid_encode_backward: dtype must be NGP_F32 or NGP_F16");
}

extern "C" int ngp_grad_total_variation(const void* inputs, const void* embeddings, void* grad,
                                        const int32_t* offsets, float weight, uint32_t B, uint32_t D, uint32_t C,
                                        uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                        int dtype, ngp_stream_t stream) {
    if (B == 0) return NGP_OK;
    cudaStream_t st = as_stream(stream);
    if (dtype == NGP_F16) { NGP_DISPATCH_DC(launch_tv, __half, inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, align_corners != 0, st) }
    if (dtype == NGP_F32) { NGP_DISPATCH_DC(launch_tv, float, inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, align_corners != 0, st) }
    return fail(NGP_EINVAL, "grad_total_variation: dtype must be NGP_F32 or NGP_F16");
}

// measurement hook: `blocks` x 256 threads x ops_per_thread reductions into table[entries] (f16x2 entries; entries % 4 == 0)
extern "C" int ngp_debug_red_probe(void* table_f16x2, uint32_t entries, uint32_t blocks, uint32_t ops_per_thread, int mode,
                                   ngp_stream_t stream) {
    if (!table_f16x2 || entries < 4 || (entries & 3u) || mode < 0 || mode > 2) return fail(NGP_EINVAL, "debug_red_probe: bad arguments");
    k_red_probe<<<blocks, 256, 0, as_stream(stream)>>>((__half*)table_f16x2, entries, ops_per_thread, mode);
    return check_launch("debug_red_probe");
}

// test hook (not part of the reference ABI): device-computed per-level scales
extern "C" int ngp_grid_level_scales(float* out_device, uint32_t L, float S, uint32_t H, ngp_stream_t stream) {
    if (L > 1024) return fail(NGP_EINVAL, "grid_level_scales: L too large");
    k_level_scales<<<1, 1024, 0, as_stream(stream)>>>(out_device, L, S, H);
    return check_launch("grid_level_scales");
}
