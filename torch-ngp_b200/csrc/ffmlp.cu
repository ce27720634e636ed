// ffmlp.cu — fully-fused bias-free fp16 MLP (64 wide) on Blackwell tcgen05 / TMEM, sm_100a.
//
// Replaces ffmlp/src/ffmlp.cu (kernel_mlp_fused :331-407, kernel_mlp_fused_backward :410-518) and
// the CUTLASS 2.8 split-K / dgrad GEMMs it calls (cutlass_matmul.h:404-488, ffmlp.cu:804-886).
//
// Semantics kept: weights are [out,in] row-major matrices back to back
// (first [64,in], then (num_layers-1) x [64,64], last [16,64]); h = act(x W0^T), ..., y = h Wout^T;
// forward_buffer[l] = post-activation output of hidden layer l; backward_buffer[j] = dL/d(pre-act)
// of hidden layer (num_layers-1-j); grad_weights in the weights' layout.
// Numerics: fp16 operands, **fp32 accumulation in TMEM** (the reference accumulates in fp16 inside
// wmma and in its split-K reduce); activations are applied to the fp32 accumulator and rounded to
// fp16 once.  Results therefore differ from the reference by its own fp16 accumulation error
// (tests bound this at 1e-3 of the output scale, the north_star tolerance).
//
// Kernel structure (one CTA = 128 threads = 128 batch rows, persistent over row tiles):
//   * every weight matrix of the network is staged ONCE per CTA into shared memory as a
//     128B-swizzled K-major UMMA operand and stays there for the CTA's lifetime,
//   * the activation tile [128 x 64] fp16 lives in shared memory (16 KB) and is overwritten in
//     place layer after layer — activations never round-trip through HBM between layers,
//   * one elected thread issues tcgen05.mma (M=128, N=64|16, K=16 per instruction), completion is
//     signalled with tcgen05.commit -> mbarrier, the 4 warps read their 32 TMEM lanes back with
//     tcgen05.ld (thread = row), apply the activation and write the next operand tile.
//   * weight gradients: a second tcgen05 kernel contracts over the batch with both operands
//     MN-major (the same row-major activation tiles, no transposes), M=64 N=64, fp32 TMEM
//     accumulators persistent across the CTA's row tiles, then one fp32 red.add per element.
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

// ================================ forward / inference ==========================================
template <bool TRAIN, uint32_t ACT>
__global__ void __launch_bounds__(128)
k_ffmlp_forward(const __half* __restrict__ inputs, const __half* __restrict__ weights,
                __half* __restrict__ forward_buffer, __half* __restrict__ outputs, const uint32_t B,
                const uint32_t in_dim, const uint32_t num_layers) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;

    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t base = align1024(smem_u32(smem_dyn));
    const uint32_t a_addr = base;
    const uint32_t w_addr = base + A_TILE_BYTES;
    const uint32_t nmat = num_layers + 1;

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<64>(&tmem_base_s);

    // stage all weight matrices (K-major B operands: row n = output neuron, 128-byte pitch)
    {
        const __half* w = weights;
        load_tile_rowmajor(w_addr, w, HID, in_dim, tid, 128);
        w += HID * in_dim;
        for (uint32_t l = 1; l < num_layers; ++l) {
            load_tile_rowmajor(w_addr + l * W_SLOT_BYTES, w, HID, HID, tid, 128);
            w += HID * HID;
        }
        load_tile_rowmajor(w_addr + num_layers * W_SLOT_BYTES, w, OUT_PAD, HID, tid, 128);
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t t_lane = tmem_base + ((warp * 32u) << 16);
    uint32_t phase = 0;

    const uint32_t ntiles = (B + TILE_M - 1) / TILE_M;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t row0 = (size_t)tile * TILE_M;
        const size_t row = row0 + tid;
        const uint32_t rows_valid = (uint32_t)((size_t)B - row0 < TILE_M ? (size_t)B - row0 : TILE_M);
        const bool row_ok = tid < rows_valid;
        // input tile -> A operand
        load_tile_rowmajor(a_addr, inputs + row0 * in_dim, TILE_M, in_dim, tid, 128, rows_valid);

        for (uint32_t l = 0; l < nmat; ++l) {
            const bool last = (l == nmat - 1);
            const uint32_t K = (l == 0) ? in_dim : HID;
            const uint32_t N = last ? OUT_PAD : HID;
            fence_async_smem();
            fence_before_sync();
            __syncthreads();
            if (warp == 0) {
                // one elected lane issues; the rest of warp 0 parks at __syncwarp (NOT in the mbarrier spin
                // loop: a spinning sibling lane can starve the issuing lane of its own warp)
                if (tid == 0) {
                    fence_after_sync();
                    issue_layer(tmem_base, a_addr, w_addr + l * W_SLOT_BYTES, N, K);
                    mma_commit(&bar);
                }
                __syncwarp();
            }
            mbar_wait(&bar, phase);
            phase ^= 1u;
            fence_after_sync();

            if (!last) {
                __half* fb = TRAIN ? forward_buffer + ((size_t)l * B + row) * HID : nullptr;
#pragma unroll
                for (uint32_t half_i = 0; half_i < 2; ++half_i) {
                    uint32_t v[32];
                    tmem_ld32(t_lane + half_i * 32, v);
                    tmem_ld_wait();
                    uint32_t p[16];
#pragma unroll
                    for (uint32_t i = 0; i < 16; ++i)
                        p[i] = pack_h2(act_fwd<ACT>(__uint_as_float(v[2 * i])), act_fwd<ACT>(__uint_as_float(v[2 * i + 1])));
#pragma unroll
                    for (uint32_t c = 0; c < 4; ++c) {
                        const uint4 q = make_uint4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
                        st_shared_v4(a_addr + sw128_off(tid, half_i * 4 + c), q);
                        if (TRAIN && row_ok) reinterpret_cast<uint4*>(fb)[half_i * 4 + c] = q;
                    }
                }
            } else {
                uint32_t v[16];
                tmem_ld16(t_lane, v);
                tmem_ld_wait();
                uint32_t p[8];
#pragma unroll
                for (uint32_t i = 0; i < 8; ++i) p[i] = pack_h2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                if (row_ok) {
                    uint4* o = reinterpret_cast<uint4*>(outputs + row * OUT_PAD);
                    o[0] = make_uint4(p[0], p[1], p[2], p[3]);
                    o[1] = make_uint4(p[4], p[5], p[6], p[7]);
                }
            }
        }
        // the next tile's input load overwrites the A tile: every warp must be done reading TMEM /
        // the last MMA must be done reading smem (it is: we waited on its commit).
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
    }

    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem_base);
}

// ================================ backward: activation gradients ===============================
// grad [B,16]; forward_buffer [num_layers,B,64]; backward_buffer [num_layers,B,64];
// grad_inputs [B,in_dim] or null.
template <uint32_t ACT>
__global__ void __launch_bounds__(128)
k_ffmlp_backward(const __half* __restrict__ grad, const __half* __restrict__ weights,
                 const __half* __restrict__ forward_buffer, __half* __restrict__ backward_buffer,
                 __half* __restrict__ grad_inputs, const uint32_t B, const uint32_t in_dim,
                 const uint32_t num_layers) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;

    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t base = align1024(smem_u32(smem_dyn));
    unsigned char* base_gen = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t a_addr = base;
    const uint32_t w_off = A_TILE_BYTES;        // offset of weight slots from `base`
    const uint32_t w_addr = base + w_off;
    const uint32_t n_hidden = num_layers - 1;   // hidden x hidden matmuls

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<64>(&tmem_base_s);

    // transposed weights, in the order the backward consumes them:
    //   slot 0            : Wout^T   tile(n = hidden, k = out 0..15)
    //   slot 1..n_hidden  : W_k^T for k = n_hidden .. 1   tile(n = in, k = out)
    //   slot n_hidden + 1 : W_0^T    tile(n = input feature, k = hidden)   (only if grad_inputs)
    {
        const __half* w0 = weights;
        const __half* wh = weights + HID * in_dim;
        const __half* wout = wh + (size_t)n_hidden * HID * HID;
        load_tile_transposed(base_gen, w_off, wout, OUT_PAD, HID, tid, 128);
        for (uint32_t j = 0; j < n_hidden; ++j)
            load_tile_transposed(base_gen, w_off + (1 + j) * W_SLOT_BYTES, wh + (size_t)(n_hidden - 1 - j) * HID * HID, HID, HID, tid, 128);
        if (grad_inputs) load_tile_transposed(base_gen, w_off + (1 + n_hidden) * W_SLOT_BYTES, w0, HID, in_dim, tid, 128);
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t t_lane = tmem_base + ((warp * 32u) << 16);
    uint32_t phase = 0;

    const uint32_t ntiles = (B + TILE_M - 1) / TILE_M;
    const uint32_t nrounds = 1 + n_hidden + (grad_inputs ? 1u : 0u);
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t row0 = (size_t)tile * TILE_M;
        const size_t row = row0 + tid;
        const uint32_t rows_valid = (uint32_t)((size_t)B - row0 < TILE_M ? (size_t)B - row0 : TILE_M);
        const bool row_ok = tid < rows_valid;
        load_tile_rowmajor(a_addr, grad + row0 * OUT_PAD, TILE_M, OUT_PAD, tid, 128, rows_valid);

        for (uint32_t r = 0; r < nrounds; ++r) {
            const bool to_inputs = grad_inputs && (r == nrounds - 1);
            const uint32_t K = (r == 0) ? OUT_PAD : HID;
            const uint32_t N = to_inputs ? in_dim : HID;
            fence_async_smem();
            fence_before_sync();
            __syncthreads();
            if (warp == 0) {
                if (tid == 0) {
                    fence_after_sync();
                    issue_layer(tmem_base, a_addr, w_addr + r * W_SLOT_BYTES, N, K);
                    mma_commit(&bar);
                }
                __syncwarp();
            }
            mbar_wait(&bar, phase);
            phase ^= 1u;
            fence_after_sync();

            if (!to_inputs) {
                // round r produces dL/d(pre-activation) of hidden layer (num_layers-1-r)
                const uint32_t layer = num_layers - 1 - r;
                const uint4* fwd = reinterpret_cast<const uint4*>(forward_buffer + ((size_t)layer * B + row) * HID);
                uint4* bb = reinterpret_cast<uint4*>(backward_buffer + ((size_t)r * B + row) * HID);
#pragma unroll
                for (uint32_t half_i = 0; half_i < 2; ++half_i) {
                    uint32_t v[32];
                    tmem_ld32(t_lane + half_i * 32, v);
                    uint4 f[4];
#pragma unroll
                    for (uint32_t c = 0; c < 4; ++c) f[c] = row_ok ? __ldg(fwd + half_i * 4 + c) : make_uint4(0, 0, 0, 0);
                    tmem_ld_wait();
                    const __half2* fh = reinterpret_cast<const __half2*>(f);
                    uint32_t p[16];
#pragma unroll
                    for (uint32_t i = 0; i < 16; ++i) {
                        const float2 ff = __half22float2(fh[i]);
                        p[i] = pack_h2(act_bwd<ACT>(__uint_as_float(v[2 * i]), ff.x), act_bwd<ACT>(__uint_as_float(v[2 * i + 1]), ff.y));
                    }
#pragma unroll
                    for (uint32_t c = 0; c < 4; ++c) {
                        const uint4 q = make_uint4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
                        st_shared_v4(a_addr + sw128_off(tid, half_i * 4 + c), q);
                        if (row_ok) bb[half_i * 4 + c] = q;
                    }
                }
            } else {
                __half* gi = grad_inputs + row * in_dim;
                for (uint32_t c0 = 0; c0 < in_dim; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_lane + c0, v);
                    tmem_ld_wait();
                    uint32_t p[8];
#pragma unroll
                    for (uint32_t i = 0; i < 8; ++i) p[i] = pack_h2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                    if (row_ok) {
                        uint4* o = reinterpret_cast<uint4*>(gi + c0);
                        o[0] = make_uint4(p[0], p[1], p[2], p[3]);
                        o[1] = make_uint4(p[4], p[5], p[6], p[7]);
                    }
                }
            }
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
    }

    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem_base);
}


// ================================ backward: fused dgrad + wgrad ================================
// One pass over the batch produces the activation gradients AND the weight gradients: every dPre / H tile is
// consumed from shared memory by BOTH the next dgrad MMA (K-major view) and the weight-gradient MMA (MN-major view of
// the same bytes), so backward_buffer never has to be written or re-read (the reference writes it, then re-reads it
// and forward_buffer in (num_layers+1) separate CUTLASS split-K GEMMs, ffmlp.cu:804-875).
//   TMEM (256 columns): [0,64) dgrad accumulator (M=128); weight-grad accumulator l (M=64,N=64 fp32, persistent over
//   the CTA's tiles) at columns 64 + 64*(l/2), lanes +16*(l%2) (two M=64 accumulators interleave in one column range).
//   smem: G0,G1 (dPre tiles, ping-pong), F0,F1 (forward-activation tiles, ping-pong), transposed weights.
//   tcgen05.mma executes in issue order, so committing [wgrad(r-1) MMAs, dgrad(r) MMAs] to one mbarrier and waiting for
//   it orders every buffer reuse.
// Supports num_layers + 1 <= 6 matmuls (NeRF: 3 and 4); wider nets fall back to the two-kernel path.
static constexpr uint32_t FUSED_MAX_MATMULS = 6;

__device__ __forceinline__ void issue_wgrad(uint32_t acc_tmem, uint32_t p_addr, uint32_t q_addr, uint32_t accumulate) {
    const uint32_t idesc = make_idesc(64, 64, 1, 1);
#pragma unroll
    for (uint32_t k = 0; k < TILE_M / 16; ++k)
        mma_f16(acc_tmem, make_desc(p_addr + k * 2048, 16384, 1024, LAYOUT_SW128),
                make_desc(q_addr + k * 2048, 16384, 1024, LAYOUT_SW128), idesc, (accumulate || k > 0) ? 1u : 0u);
}

template <uint32_t ACT>
__global__ void __launch_bounds__(128)
k_ffmlp_backward_fused(const __half* __restrict__ grad, const __half* __restrict__ inputs,
                       const __half* __restrict__ weights, const __half* __restrict__ forward_buffer,
                       __half* __restrict__ backward_buffer, __half* __restrict__ grad_inputs,
                       float* __restrict__ wgrad_ws, const uint32_t B, const uint32_t in_dim, const uint32_t num_layers) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t base = align1024(smem_u32(smem_dyn));
    unsigned char* base_gen = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t g_addr[2] = {base, base + A_TILE_BYTES};
    const uint32_t f_addr[2] = {base + 2 * A_TILE_BYTES, base + 3 * A_TILE_BYTES};
    const uint32_t w_off = 4 * A_TILE_BYTES;
    const uint32_t w_addr = base + w_off;
    const uint32_t n_hidden = num_layers - 1;
    const uint32_t nmat = num_layers + 1;

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<256>(&tmem_base_s);
    {   // transposed weights in consumption order (see k_ffmlp_backward)
        const __half* w0 = weights;
        const __half* wh = weights + HID * in_dim;
        const __half* wout = wh + (size_t)n_hidden * HID * HID;
        load_tile_transposed(base_gen, w_off, wout, OUT_PAD, HID, tid, 128);
        for (uint32_t j = 0; j < n_hidden; ++j)
            load_tile_transposed(base_gen, w_off + (1 + j) * W_SLOT_BYTES, wh + (size_t)(n_hidden - 1 - j) * HID * HID, HID, HID, tid, 128);
        if (grad_inputs) load_tile_transposed(base_gen, w_off + (1 + n_hidden) * W_SLOT_BYTES, w0, HID, in_dim, tid, 128);
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t t_lane = tmem_base + ((warp * 32u) << 16);
    // weight-gradient accumulator of matmul l (l = 0 .. nmat-1, weights' order)
    auto acc_addr = [&](uint32_t l) { return tmem_base + 64u + 64u * (l >> 1) + ((16u * (l & 1u)) << 16); };
    uint32_t phase = 0;
    bool first_tile = true;

    const uint32_t ntiles = (B + TILE_M - 1) / TILE_M;
    const uint32_t nrounds = 1 + n_hidden + (grad_inputs ? 1u : 0u);
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t row0 = (size_t)tile * TILE_M;
        const size_t row = row0 + tid;
        const uint32_t rows_valid = (uint32_t)((size_t)B - row0 < TILE_M ? (size_t)B - row0 : TILE_M);
        const bool row_ok = tid < rows_valid;
        // dL/dy -> G1 (K-major A of round 0 and, zero-padded to 64 columns, MN-major operand of the output-layer wgrad)
        load_tile_rowmajor(g_addr[1], grad + row0 * OUT_PAD, TILE_M, OUT_PAD, tid, 128, rows_valid);
        zero_tile_cols(g_addr[1], TILE_M, OUT_PAD >> 3, tid, 128);

        // round r consumes A = (r == 0 ? G1 : G[(r-1)&1]) and produces dPre of hidden layer (num_layers-1-r) in G[r&1],
        // with the matching forward activations H in F[r&1].
        for (uint32_t r = 0; r < nrounds; ++r) {
            const bool to_inputs = grad_inputs && (r == nrounds - 1);
            const uint32_t K = (r == 0) ? OUT_PAD : HID;
            const uint32_t N = to_inputs ? in_dim : HID;
            const uint32_t a_in = (r == 0) ? g_addr[1] : g_addr[(r - 1) & 1u];
            // prefetch this round's forward activations (ReLU mask + wgrad operand) while the MMAs run
            uint4 f[8];
            if (!to_inputs) {
                const uint4* fwd = reinterpret_cast<const uint4*>(forward_buffer + ((size_t)(num_layers - 1 - r) * B + row) * HID);
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) f[c] = row_ok ? __ldg(fwd + c) : make_uint4(0, 0, 0, 0);
            }
            fence_async_smem();
            fence_before_sync();
            __syncthreads();
            if (warp == 0) {
                if (tid == 0) {
                    fence_after_sync();
                    // weight gradient of the matmul consumed in the PREVIOUS round: its dPre and H tiles are complete now
                    if (r == 1) {          // output layer: P = dY (G1, 16 valid columns), Q = H_{nl-1} (F0)
                        issue_wgrad(acc_addr(nmat - 1), g_addr[1], f_addr[0], first_tile ? 0u : 1u);
                    } else if (r >= 2) {   // hidden matmul (num_layers+1-r): P = dPre (G[(r-2)&1]), Q = H (F[(r-1)&1])
                        issue_wgrad(acc_addr(num_layers + 1 - r), g_addr[(r - 2) & 1u], f_addr[(r - 1) & 1u], first_tile ? 0u : 1u);
                    }
                    issue_layer(tmem_base, a_in, w_addr + r * W_SLOT_BYTES, N, K);
                    mma_commit(&bar);
                }
                __syncwarp();
            }
            mbar_wait(&bar, phase);
            phase ^= 1u;
            fence_after_sync();

            if (!to_inputs) {
                const uint32_t gw = g_addr[r & 1u], fw = f_addr[r & 1u];
                uint4* bb = backward_buffer ? reinterpret_cast<uint4*>(backward_buffer + ((size_t)r * B + row) * HID) : nullptr;
#pragma unroll
                for (uint32_t half_i = 0; half_i < 2; ++half_i) {
                    uint32_t v[32];
                    tmem_ld32(t_lane + half_i * 32, v);
                    tmem_ld_wait();
                    const __half2* fh = reinterpret_cast<const __half2*>(&f[half_i * 4]);
                    uint32_t p[16];
#pragma unroll
                    for (uint32_t i = 0; i < 16; ++i) {
                        const float2 ff = __half22float2(fh[i]);
                        p[i] = pack_h2(act_bwd<ACT>(__uint_as_float(v[2 * i]), ff.x), act_bwd<ACT>(__uint_as_float(v[2 * i + 1]), ff.y));
                    }
#pragma unroll
                    for (uint32_t c = 0; c < 4; ++c) {
                        const uint4 q = make_uint4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
                        st_shared_v4(gw + sw128_off(tid, half_i * 4 + c), q);
                        st_shared_v4(fw + sw128_off(tid, half_i * 4 + c), f[half_i * 4 + c]);
                        if (bb && row_ok) bb[half_i * 4 + c] = q;
                    }
                }
            } else {
                __half* gi = grad_inputs + row * in_dim;
                for (uint32_t c0 = 0; c0 < in_dim; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_lane + c0, v);
                    tmem_ld_wait();
                    uint32_t p[8];
#pragma unroll
                    for (uint32_t i = 0; i < 8; ++i) p[i] = pack_h2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                    if (row_ok) {
                        uint4* o = reinterpret_cast<uint4*>(gi + c0);
                        o[0] = make_uint4(p[0], p[1], p[2], p[3]);
                        o[1] = make_uint4(p[4], p[5], p[6], p[7]);
                    }
                }
            }
        }

        // tail of the tile: wgrad of the last hidden matmul handled above (if any) and of matmul 0 (P = dPre_0, Q = X).
        // dPre_0 was produced in round n_hidden -> G[n_hidden & 1]; H_0 is in F[n_hidden & 1]; X goes to the other F.
        {
            const uint32_t xq = f_addr[(n_hidden + 1) & 1u];
            load_tile_rowmajor(xq, inputs + row0 * in_dim, TILE_M, in_dim, tid, 128, rows_valid);
            zero_tile_cols(xq, TILE_M, in_dim >> 3, tid, 128);
            fence_async_smem();
            fence_before_sync();
            __syncthreads();
            if (warp == 0) {
                if (tid == 0) {
                    fence_after_sync();
                    if (!grad_inputs) {
                        // the wgrad that the (absent) input round would have issued: matmul 1 (or the output layer when n_hidden == 0)
                        const uint32_t r = nrounds;   // == 1 + n_hidden
                        if (r == 1) issue_wgrad(acc_addr(nmat - 1), g_addr[1], f_addr[0], first_tile ? 0u : 1u);
                        else issue_wgrad(acc_addr(num_layers + 1 - r), g_addr[(r - 2) & 1u], f_addr[(r - 1) & 1u], first_tile ? 0u : 1u);
                    }
                    issue_wgrad(acc_addr(0), g_addr[n_hidden & 1u], xq, first_tile ? 0u : 1u);
                    mma_commit(&bar);
                }
                __syncwarp();
            }
            mbar_wait(&bar, phase);
            phase ^= 1u;
            fence_after_sync();
        }
        first_tile = false;
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
    }

    // flush the CTA's weight-gradient accumulators (fp32) into the workspace laid out like `weights`
    if (!first_tile) {
        for (uint32_t lp = 0; lp < (nmat + 1) / 2; ++lp) {
            const uint32_t l = lp * 2 + (lane >> 4);                  // lanes 0-15: even matmul, 16-31: odd matmul
            const uint32_t m = warp * 16 + (lane & 15u);
            uint32_t Mv, Nv, ws_off;
            if (l == 0) { Mv = HID; Nv = in_dim; ws_off = 0; }
            else if (l < nmat - 1) { Mv = HID; Nv = HID; ws_off = HID * in_dim + (l - 1) * HID * HID; }
            else { Mv = OUT_PAD; Nv = HID; ws_off = HID * in_dim + (num_layers - 1) * HID * HID; }
#pragma unroll
            for (uint32_t half_i = 0; half_i < 2; ++half_i) {
                uint32_t v[32];
                tmem_ld32(t_lane + 64u + 64u * lp + half_i * 32, v);
                tmem_ld_wait();
                if (l < nmat && m < Mv) {
                    float* dst = wgrad_ws + ws_off + (size_t)m * Nv;
#pragma unroll
                    for (uint32_t i = 0; i < 32; ++i) {
                        const uint32_t n = half_i * 32 + i;
                        if (n < Nv) atomicAdd(dst + n, __uint_as_float(v[i]));
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem_base);
}

// ================================ backward: weight gradients ===================================
// dW_l[m][n] = sum_rows P_l[row][m] * Q_l[row][n]   (contraction over the batch)
//   l = 0            : P = dPre_0 (backward_buffer[num_layers-1]),    Q = inputs (in_dim wide)
//   1 <= l < nmat-1  : P = dPre_l (backward_buffer[num_layers-1-l]),  Q = forward_buffer[l-1]
//   l = nmat-1       : P = dL/dy  (16 wide),                          Q = forward_buffer[num_layers-1]
// grid = (nmat, nsplit); each CTA accumulates its row tiles in one fp32 TMEM accumulator (M=64,N=64)
// and finally red.adds the valid [Mv x Nv] block into the fp32 workspace (layout of `weights`).
static constexpr uint32_t WG_STAGE_BYTES = 2 * A_TILE_BYTES;   // P tile + Q tile

__global__ void __launch_bounds__(128)
k_ffmlp_wgrad(const __half* __restrict__ grad, const __half* __restrict__ inputs,
              const __half* __restrict__ forward_buffer, const __half* __restrict__ backward_buffer,
              float* __restrict__ wgrad_ws, const uint32_t B, const uint32_t in_dim, const uint32_t num_layers) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t base = align1024(smem_u32(smem_dyn));
    const uint32_t nmat = num_layers + 1;
    const uint32_t l = blockIdx.x;

    const __half* P; const __half* Q;
    uint32_t Pw, Qw, Mv, Nv, ws_off;
    if (l == 0) {
        P = backward_buffer + (size_t)(num_layers - 1) * B * HID; Pw = HID;
        Q = inputs; Qw = in_dim; Mv = HID; Nv = in_dim; ws_off = 0;
    } else if (l < nmat - 1) {
        P = backward_buffer + (size_t)(num_layers - 1 - l) * B * HID; Pw = HID;
        Q = forward_buffer + (size_t)(l - 1) * B * HID; Qw = HID; Mv = HID; Nv = HID;
        ws_off = HID * in_dim + (l - 1) * HID * HID;
    } else {
        P = grad; Pw = OUT_PAD;
        Q = forward_buffer + (size_t)(num_layers - 1) * B * HID; Qw = HID; Mv = OUT_PAD; Nv = HID;
        ws_off = HID * in_dim + (num_layers - 1) * HID * HID;
    }

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<64>(&tmem_base_s);
    // zero the padding columns of both stages once (they are never overwritten by the row loads)
    for (uint32_t s = 0; s < 2; ++s) {
        zero_tile_cols(base + s * WG_STAGE_BYTES, TILE_M, Pw >> 3, tid, 128);
        zero_tile_cols(base + s * WG_STAGE_BYTES + A_TILE_BYTES, TILE_M, Qw >> 3, tid, 128);
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    const uint32_t idesc = make_idesc(64, 64, 1, 1);   // both operands MN-major
    const uint32_t ntiles = (B + TILE_M - 1) / TILE_M;
    uint32_t ph[2] = {0, 0};
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.y; tile < ntiles; tile += gridDim.y, ++it) {
        const uint32_t s = it & 1u;
        const uint32_t p_addr = base + s * WG_STAGE_BYTES, q_addr = p_addr + A_TILE_BYTES;
        if (it >= 2) { mbar_wait(&bars[s], ph[s]); ph[s] ^= 1u; }   // MMAs that read this stage are done
        const size_t row0 = (size_t)tile * TILE_M;
        const uint32_t rows_valid = (uint32_t)((size_t)B - row0 < TILE_M ? (size_t)B - row0 : TILE_M);
        load_tile_rowmajor(p_addr, P + row0 * Pw, TILE_M, Pw, tid, 128, rows_valid);
        load_tile_rowmajor(q_addr, Q + row0 * Qw, TILE_M, Qw, tid, 128, rows_valid);
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (warp == 0) {
          if (tid == 0) {
            fence_after_sync();
#pragma unroll
            for (uint32_t k = 0; k < TILE_M / 16; ++k) {
                // MN-major SW128: 16 batch rows (= MMA K) advance by 16 * 128 B; 8-row groups 1024 B apart
                const uint64_t ad = make_desc(p_addr + k * 2048, 16384, 1024, LAYOUT_SW128);
                const uint64_t bd = make_desc(q_addr + k * 2048, 16384, 1024, LAYOUT_SW128);
                mma_f16(tmem_base, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
            mma_commit(&bars[s]);
          }
          __syncwarp();
        }
    }
    // drain: wait for the last commit of each stage that was used
    const uint32_t n_it = it;
    if (n_it >= 1) { const uint32_t s = (n_it - 1) & 1u; mbar_wait(&bars[s], ph[s]); ph[s] ^= 1u; }
    if (n_it >= 2) { const uint32_t s = (n_it - 2) & 1u; mbar_wait(&bars[s], ph[s]); ph[s] ^= 1u; }
    fence_after_sync();

    if (n_it > 0) {
        // M=64 accumulator: row m lives on TMEM lane (m % 16) + 32 * (m / 16) -> warp w, lanes 0..15
        const uint32_t t_lane = tmem_base + ((warp * 32u) << 16);
        const uint32_t m = warp * 16 + lane;
#pragma unroll
        for (uint32_t half_i = 0; half_i < 2; ++half_i) {
            uint32_t v[32];
            tmem_ld32(t_lane + half_i * 32, v);
            tmem_ld_wait();
            if (lane < 16 && m < Mv) {
                float* dst = wgrad_ws + ws_off + (size_t)m * Nv;
#pragma unroll
                for (uint32_t i = 0; i < 32; ++i) {
                    const uint32_t n = half_i * 32 + i;
                    if (n < Nv) atomicAdd(dst + n, __uint_as_float(v[i]));
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem_base);
}

__global__ void k_ffmlp_wgrad_finalize(const float* __restrict__ ws, __half* __restrict__ grad_weights, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) grad_weights[i] = __float2half_rn(ws[i]);
}

// ================================ debug probe (tests only) =====================================
// mode 0: D[128x64] = A[128x64] * Bm[64x64]^T  (K-major operands, M=128)
// mode 1: D[64x64]  = A[128x64]^T * Bm[128x64] (MN-major operands, M=64, contraction over 128 rows)
__global__ void __launch_bounds__(128)
k_umma_probe(const __half* __restrict__ A, const __half* __restrict__ Bm, float* __restrict__ D, int mode) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t base = align1024(smem_u32(smem_dyn));
    const uint32_t a_addr = base, b_addr = base + A_TILE_BYTES;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<64>(&tmem_base_s);
    load_tile_rowmajor(a_addr, A, TILE_M, HID, tid, 128);
    load_tile_rowmajor(b_addr, Bm, mode == 0 ? HID : TILE_M, HID, tid, 128);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0) {
      if (tid == 0) {
        if (mode == 0) {
            issue_layer(tmem_base, a_addr, b_addr, HID, HID);
        } else {
            const uint32_t idesc = make_idesc(64, 64, 1, 1);
            for (uint32_t k = 0; k < TILE_M / 16; ++k)
                mma_f16(tmem_base, make_desc(a_addr + k * 2048, 16384, 1024, LAYOUT_SW128),
                        make_desc(b_addr + k * 2048, 16384, 1024, LAYOUT_SW128), idesc, k > 0 ? 1u : 0u);
        }
        mma_commit(&bar);
      }
      __syncwarp();
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    const uint32_t t_lane = tmem_base + ((warp * 32u) << 16);
    for (uint32_t half_i = 0; half_i < 2; ++half_i) {
        uint32_t v[32];
        tmem_ld32(t_lane + half_i * 32, v);
        tmem_ld_wait();
        if (mode == 0) {
            for (uint32_t i = 0; i < 32; ++i) D[(size_t)tid * 64 + half_i * 32 + i] = __uint_as_float(v[i]);
        } else if (lane < 16) {
            const uint32_t m = warp * 16 + lane;
            for (uint32_t i = 0; i < 32; ++i) D[(size_t)m * 64 + half_i * 32 + i] = __uint_as_float(v[i]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem_base);
}

// ---- host side -------------------------------------------------------------------------------
static int check_cfg(const char* who, uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                     uint32_t num_layers, uint32_t output_activation) {
    if (hidden_dim != HID) return fail(NGP_EUNSUPPORTED, "%s: this build supports hidden_dim == 64 (got %u)", who, hidden_dim);
    if (input_dim == 0 || input_dim % 16 != 0 || input_dim > 64)
        return fail(NGP_EUNSUPPORTED, "%s: input_dim must be 16, 32, 48 or 64 (got %u)", who, input_dim);
    if (output_dim != OUT_PAD) return fail(NGP_EUNSUPPORTED, "%s: padded output_dim must be 16 (got %u)", who, output_dim);
    if (num_layers < 2 || num_layers + 1 > MAX_MATMULS) return fail(NGP_EINVAL, "%s: num_layers must be in [2, 8] (got %u)", who, num_layers);
    if (output_activation != ACT_NONE) return fail(NGP_EUNSUPPORTED, "%s: output activation is not supported (ffmlp.py:108)", who);
    return NGP_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes, const char* who) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(NGP_ECUDA, "%s: insufficient shared memory (%zu bytes): %s", who, bytes, cudaGetErrorString(e));
    return NGP_OK;
}

static uint32_t persistent_grid(uint32_t ntiles, uint32_t ctas_per_sm) {
    const uint32_t cap = (uint32_t)sm_count() * ctas_per_sm;
    return ntiles < cap ? ntiles : cap;
}

}  // namespace ngp

using namespace ngp;

#define NGP_DISPATCH_ACT(ACTV, ...)                         \
    switch (ACTV) {                                         \
        case ACT_RELU: { constexpr uint32_t A = ACT_RELU; __VA_ARGS__; } break;             \
        case ACT_EXP: { constexpr uint32_t A = ACT_EXP; __VA_ARGS__; } break;               \
        case ACT_SINE: { constexpr uint32_t A = ACT_SINE; __VA_ARGS__; } break;             \
        case ACT_SIGMOID: { constexpr uint32_t A = ACT_SIGMOID; __VA_ARGS__; } break;       \
        case ACT_SQUAREPLUS: { constexpr uint32_t A = ACT_SQUAREPLUS; __VA_ARGS__; } break; \
        case ACT_SOFTPLUS: { constexpr uint32_t A = ACT_SOFTPLUS; __VA_ARGS__; } break;     \
        default: { constexpr uint32_t A = ACT_NONE; __VA_ARGS__; } break;                   \
    }

template <bool TRAIN>
static int launch_forward(const void* inputs, const void* weights, uint32_t B, uint32_t input_dim, uint32_t num_layers,
                          uint32_t activation, void* forward_buffer, void* outputs, cudaStream_t st, const char* who) {
    const size_t smem = 1024 + A_TILE_BYTES + (size_t)(num_layers + 1) * W_SLOT_BYTES;
    const uint32_t grid = persistent_grid((B + TILE_M - 1) / TILE_M, 4);
    int rc = NGP_OK;
    NGP_DISPATCH_ACT(activation,
        rc = set_smem(k_ffmlp_forward<TRAIN, A>, smem, who);
        if (rc == NGP_OK)
            k_ffmlp_forward<TRAIN, A><<<grid, 128, smem, st>>>((const __half*)inputs, (const __half*)weights,
                                                                (__half*)forward_buffer, (__half*)outputs, B, input_dim, num_layers))
    if (rc) return rc;
    return check_launch(who);
}

extern "C" int ngp_ffmlp_forward(const void* inputs, const void* weights, uint32_t B, uint32_t input_dim,
                                 uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                                 uint32_t output_activation, void* forward_buffer, void* outputs, ngp_stream_t stream) {
    if (B == 0) return NGP_OK;
    int rc = check_cfg("ffmlp_forward", B, input_dim, output_dim, hidden_dim, num_layers, output_activation);
    if (rc) return rc;
    if (!forward_buffer) return fail(NGP_EINVAL, "ffmlp_forward: forward_buffer is required (use ffmlp_inference otherwise)");
    return launch_forward<true>(inputs, weights, B, input_dim, num_layers, activation, forward_buffer, outputs, as_stream(stream), "ffmlp_forward");
}

extern "C" int ngp_ffmlp_inference(const void* inputs, const void* weights, uint32_t B, uint32_t input_dim,
                                   uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                                   uint32_t output_activation, void* inference_buffer, void* outputs,
                                   ngp_stream_t stream) {
    (void)inference_buffer;
    if (B == 0) return NGP_OK;
    int rc = check_cfg("ffmlp_inference", B, input_dim, output_dim, hidden_dim, num_layers, output_activation);
    if (rc) return rc;
    return launch_forward<false>(inputs, weights, B, input_dim, num_layers, activation, nullptr, outputs, as_stream(stream), "ffmlp_inference");
}

extern "C" size_t ngp_ffmlp_backward_workspace_bytes(uint32_t B, uint32_t input_dim, uint32_t output_dim,
                                                     uint32_t hidden_dim, uint32_t num_layers) {
    (void)B;
    return sizeof(float) * (size_t)hidden_dim * (input_dim + (size_t)hidden_dim * (num_layers - 1) + output_dim);
}

// backward_buffer may be NULL: the fused kernel then keeps dL/d(pre-activation) on chip only.
extern "C" int ngp_ffmlp_backward(const void* grad, const void* inputs, const void* weights, const void* forward_buffer,
                                  uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                                  uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                                  int calc_grad_inputs, void* backward_buffer, void* grad_inputs, void* grad_weights,
                                  void* workspace, size_t workspace_bytes, ngp_stream_t stream) {
    int rc = check_cfg("ffmlp_backward", B, input_dim, output_dim, hidden_dim, num_layers, output_activation);
    if (rc) return rc;
    const size_t need = ngp_ffmlp_backward_workspace_bytes(B, input_dim, output_dim, hidden_dim, num_layers);
    if (!workspace || workspace_bytes < need) return fail(NGP_EINVAL, "ffmlp_backward: workspace too small (%zu < %zu)", workspace_bytes, need);
    if (calc_grad_inputs && !grad_inputs) return fail(NGP_EINVAL, "ffmlp_backward: grad_inputs is null");
    cudaStream_t st = as_stream(stream);
    const uint32_t n_params = (uint32_t)(need / sizeof(float));
    if (cudaMemsetAsync(workspace, 0, need, st) != cudaSuccess) return fail(NGP_ECUDA, "ffmlp_backward: memset failed");
    const uint32_t nmat = num_layers + 1;
    __half* gi = calc_grad_inputs ? (__half*)grad_inputs : nullptr;
    if (B > 0 && nmat <= FUSED_MAX_MATMULS) {
        const uint32_t nslots = 1 + (num_layers - 1) + (calc_grad_inputs ? 1 : 0);
        const size_t smem = 1024 + 4 * (size_t)A_TILE_BYTES + (size_t)nslots * W_SLOT_BYTES;
        const uint32_t grid = persistent_grid((B + TILE_M - 1) / TILE_M, 2);
        NGP_DISPATCH_ACT(activation,
            rc = set_smem(k_ffmlp_backward_fused<A>, smem, "ffmlp_backward");
            if (rc == NGP_OK)
                k_ffmlp_backward_fused<A><<<grid, 128, smem, st>>>((const __half*)grad, (const __half*)inputs, (const __half*)weights,
                                                                   (const __half*)forward_buffer, (__half*)backward_buffer, gi,
                                                                   (float*)workspace, B, input_dim, num_layers))
        if (rc) return rc;
        rc = check_launch("ffmlp_backward");
        if (rc) return rc;
    } else if (B > 0) {
        if (!backward_buffer) return fail(NGP_EINVAL, "ffmlp_backward: backward_buffer is required for num_layers > 5");
        const uint32_t nslots = 1 + (num_layers - 1) + (calc_grad_inputs ? 1 : 0);
        const size_t smem = 1024 + A_TILE_BYTES + (size_t)nslots * W_SLOT_BYTES;
        const uint32_t grid = persistent_grid((B + TILE_M - 1) / TILE_M, 4);
        NGP_DISPATCH_ACT(activation,
            rc = set_smem(k_ffmlp_backward<A>, smem, "ffmlp_backward");
            if (rc == NGP_OK)
                k_ffmlp_backward<A><<<grid, 128, smem, st>>>((const __half*)grad, (const __half*)weights, (const __half*)forward_buffer,
                                                             (__half*)backward_buffer, gi, B, input_dim, num_layers))
        if (rc) return rc;
        rc = check_launch("ffmlp_backward");
        if (rc) return rc;

        const size_t smem_w = 1024 + 2 * (size_t)WG_STAGE_BYTES;
        rc = set_smem(k_ffmlp_wgrad, smem_w, "ffmlp_backward(wgrad)");
        if (rc) return rc;
        const uint32_t ntiles = (B + TILE_M - 1) / TILE_M;
        uint32_t nsplit = (uint32_t)sm_count() * 2 / nmat;
        if (nsplit < 1) nsplit = 1;
        if (nsplit > ntiles) nsplit = ntiles;
        k_ffmlp_wgrad<<<dim3(nmat, nsplit), 128, smem_w, st>>>((const __half*)grad, (const __half*)inputs,
                                                               (const __half*)forward_buffer, (const __half*)backward_buffer,
                                                               (float*)workspace, B, input_dim, num_layers);
        rc = check_launch("ffmlp_backward(wgrad)");
        if (rc) return rc;
    }
    k_ffmlp_wgrad_finalize<<<div_up(n_params, 256u), 256, 0, st>>>((const float*)workspace, (__half*)grad_weights, n_params);
    return check_launch("ffmlp_backward(finalize)");
}

extern "C" int ngp_ffmlp_allocate_splitk(size_t size) { (void)size; return NGP_OK; }
extern "C" int ngp_ffmlp_free_splitk(void) { return NGP_OK; }

// test hook, not part of the reference ABI
extern "C" int ngp_debug_umma(const void* A, const void* Bm, float* D, int mode, ngp_stream_t stream) {
    const size_t smem = 1024 + 2 * (size_t)A_TILE_BYTES;
    int rc = set_smem(k_umma_probe, smem, "debug_umma");
    if (rc) return rc;
    k_umma_probe<<<1, 128, smem, as_stream(stream)>>>((const __half*)A, (const __half*)Bm, D, mode);
    return check_launch("debug_umma");
}
