// mlp_common.cuh — tcgen05 building blocks shared by the FFMLP kernels (ffmlp.cu) and the fused
// encoder->MLP / SH->MLP kernels (fused.cu).
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace ngp {
using namespace umma;

static constexpr uint32_t TILE_M = 128;       // batch rows per CTA tile
static constexpr uint32_t HID = 64;           // hidden width supported by this build
static constexpr uint32_t OUT_PAD = 16;       // padded output width (ffmlp.py:118)
static constexpr uint32_t MAX_MATMULS = 9;    // num_layers + 1 <= 9
static constexpr uint32_t A_TILE_BYTES = TILE_M * 128;   // 16 KB
static constexpr uint32_t W_SLOT_BYTES = HID * 128;      // 8 KB per weight matrix slot
static constexpr float K_ACT = 10.0f;         // reference utils.h: squareplus / softplus sharpness

enum Act : uint32_t { ACT_RELU = 0, ACT_EXP = 1, ACT_SINE = 2, ACT_SIGMOID = 3, ACT_SQUAREPLUS = 4, ACT_SOFTPLUS = 5, ACT_NONE = 6 };

// Activations are compile-time template parameters: a runtime switch inlined 64x per thread per layer
// blew the kernel up to ~570 KB of SASS and made it instruction-fetch bound (ncu r1: 65% of stall samples
// "no_instructions", tensor pipe 0.9%).
template <uint32_t A>
__device__ __forceinline__ float act_fwd(float x) {
    if constexpr (A == ACT_RELU) return fmaxf(x, 0.f);
    else if constexpr (A == ACT_EXP) return expf(x);
    else if constexpr (A == ACT_SINE) return sinf(x);
    else if constexpr (A == ACT_SIGMOID) return 1.0f / (1.0f + expf(-x));
    else if constexpr (A == ACT_SQUAREPLUS) { const float t = x * K_ACT; return 0.5f * (t + sqrtf(t * t + 4)) / K_ACT; }
    else if constexpr (A == ACT_SOFTPLUS) return logf(expf(x * K_ACT) + 1.0f) / K_ACT;
    else return x;
}
// dL/dpre = g * act'(.) expressed through the stored post-activation value f (reference
// utils.h warp_activation_backward)
template <uint32_t A>
__device__ __forceinline__ float act_bwd(float g, float f) {
    if constexpr (A == ACT_RELU) return f > 0.f ? g : 0.f;
    else if constexpr (A == ACT_EXP) return g * f;
    else if constexpr (A == ACT_SIGMOID) return g * (f * (1.0f - f));
    else if constexpr (A == ACT_SQUAREPLUS) { const float y = f * K_ACT; return g * (y * y / (y * y + 1)); }
    else if constexpr (A == ACT_SOFTPLUS) return g * (1.0f - expf(-f * K_ACT));
    else return g;   // None; Sine: the reference leaves the gradient unchanged (needs pre-activations)
}

__device__ __forceinline__ uint32_t align1024(uint32_t a) { return (a + 1023u) & ~1023u; }

// copy a row-major [rows x cols] fp16 matrix (cols % 8 == 0) from global into a swizzled tile,
// 16-byte chunks, coalesced along the source rows.
// rows >= rows_valid (ragged last batch tile) are filled with zeros.
__device__ __forceinline__ void load_tile_rowmajor(uint32_t tile_addr, const __half* __restrict__ src, uint32_t rows,
                                                   uint32_t cols, uint32_t tid, uint32_t nthr, uint32_t rows_valid = 0xffffffffu) {
    const uint32_t cpr = cols >> 3;
    const uint32_t total = rows * cpr;
    for (uint32_t g = tid; g < total; g += nthr) {
        const uint32_t r = g / cpr, c = g - r * cpr;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < rows_valid) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * cols) + c);
        st_shared_v4(tile_addr + sw128_off(r, c), v);
    }
}
// asynchronous variant: cp.async (LDGSTS) 16-byte copies, zero-filled for rows >= rows_valid; the caller commits the
// group, later waits (cp_async_wait_all) and barriers before anyone reads the tile.
__device__ __forceinline__ void load_tile_rowmajor_async(uint32_t tile_addr, const __half* __restrict__ src, uint32_t rows,
                                                         uint32_t cols, uint32_t tid, uint32_t nthr, uint32_t rows_valid) {
    const uint32_t cpr = cols >> 3;
    const uint32_t total = rows * cpr;
    for (uint32_t g = tid; g < total; g += nthr) {
        const uint32_t r = g / cpr, c = g - r * cpr;
        const bool ok = r < rows_valid;
        const void* gp = reinterpret_cast<const uint4*>(src + (size_t)(ok ? r : 0) * cols) + c;
        const uint32_t nbytes = ok ? 16u : 0u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile_addr + sw128_off(r, c)), "l"(gp), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// same copy without the commit: the caller groups several tiles / commits (possibly empty) groups itself
__device__ __forceinline__ void load_tile_rowmajor_async_nocommit(uint32_t tile_addr, const __half* __restrict__ src, uint32_t rows,
                                                                  uint32_t cols, uint32_t tid, uint32_t nthr, uint32_t rows_valid) {
    const uint32_t cpr = cols >> 3;
    const uint32_t total = rows * cpr;
    for (uint32_t g = tid; g < total; g += nthr) {
        const uint32_t r = g / cpr, c = g - r * cpr;
        const bool ok = r < rows_valid;
        const void* gp = reinterpret_cast<const uint4*>(src + (size_t)(ok ? r : 0) * cols) + c;
        const uint32_t nbytes = ok ? 16u : 0u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile_addr + sw128_off(r, c)), "l"(gp), "r"(nbytes) : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `n` of this thread's most recent cp.async groups are still in flight (n is a run-time value <= 7)
__device__ __forceinline__ void cp_async_wait_pending(uint32_t n) {
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
}
// 128-thread named barrier (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync_128(uint32_t id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// zero-fill chunks [c0, 8) of every row of a tile
__device__ __forceinline__ void zero_tile_cols(uint32_t tile_addr, uint32_t rows, uint32_t c0, uint32_t tid, uint32_t nthr) {
    const uint32_t cpr = 8 - c0;
    if (cpr == 0) return;
    const uint32_t total = rows * cpr;
    for (uint32_t g = tid; g < total; g += nthr) {
        const uint32_t r = g / cpr, c = c0 + (g - r * cpr);
        st_shared_v4(tile_addr + sw128_off(r, c), make_uint4(0, 0, 0, 0));
    }
}
// store the TRANSPOSE of a row-major [k_rows x n_cols] matrix: tile(n, k) = src[k][n]
__device__ __forceinline__ void load_tile_transposed(unsigned char* smem_generic, uint32_t tile_off,
                                                     const __half* __restrict__ src, uint32_t k_rows, uint32_t n_cols,
                                                     uint32_t tid, uint32_t nthr) {
    const uint32_t total = k_rows * n_cols;
    for (uint32_t g = tid; g < total; g += nthr) {
        const uint32_t k = g / n_cols, n = g - k * n_cols;
        const __half v = src[g];
        *reinterpret_cast<__half*>(smem_generic + tile_off + sw128_off(n, k >> 3) + ((k & 7u) << 1)) = v;
    }
}

// issue one layer: D[128 x N] = A[128 x K] * W^T, K-major SW128 operands
__device__ __forceinline__ void issue_layer(uint32_t d_tmem, uint32_t a_addr, uint32_t w_addr, uint32_t N, uint32_t K) {
    const uint32_t idesc = make_idesc(TILE_M, N, 0, 0);
    for (uint32_t k = 0; k < K; k += 16) {
        const uint64_t ad = make_desc(a_addr + k * 2, 16, 1024, LAYOUT_SW128);
        const uint64_t bd = make_desc(w_addr + k * 2, 16, 1024, LAYOUT_SW128);
        mma_f16(d_tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
}

__device__ __forceinline__ void issue_wgrad(uint32_t acc_tmem, uint32_t p_addr, uint32_t q_addr, uint32_t accumulate) {
    const uint32_t idesc = make_idesc(64, 64, 1, 1);
#pragma unroll
    for (uint32_t k = 0; k < TILE_M / 16; ++k)
        mma_f16(acc_tmem, make_desc(p_addr + k * 2048, 16384, 1024, LAYOUT_SW128),
                make_desc(q_addr + k * 2048, 16384, 1024, LAYOUT_SW128), idesc, (accumulate || k > 0) ? 1u : 0u);
}


}  // namespace ngp
