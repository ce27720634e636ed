// density_grid.cu — occupancy-grid maintenance on the device.
//
// The reference keeps the marcher's occupancy grid up to date from Python (nerf/renderer.py): mark_untrained_grid
// (:380-442) is a 5-level loop of meshgrid / batched matmul / mask ops, update_extra_state (:445-538) builds the sample
// positions with a dozen elementwise torch ops per cascade, calls torch.nonzero (a host sync), scatters into a
// temporary grid, does a masked EMA-max, a mean with .item() (another sync) and finally packbits.  Here each of those
// stages is one kernel over the Morton-ordered grid [C, H^3] (the layout the marcher reads, raymarching.cu:279-288,378-379),
// the threshold stays on the device, and nothing synchronises with the host:
//
//   ngp_density_grid_mark_untrained   cells seen by no training camera -> -1                 (renderer.py:380-442)
//   ngp_density_grid_occupied         ordered list of cells with density > 0 (= torch.nonzero)  (renderer.py:495)
//   ngp_density_grid_sample_full      jittered position of every cell, Morton order           (renderer.py:456-483)
//   ngp_density_grid_sample_partial   H^3/4 uniform + H^3/4 occupied cells per cascade         (renderer.py:487-515)
//   ngp_density_grid_update           scatter -> EMA-max -> mean -> threshold -> bitfield     (renderer.py:521-530)
//
// The density query between "sample" and "update" is the fused encoder+MLP kernel ngp_field_sigma_forward (or any
// density function of the caller).  Arithmetic follows the torch expressions of the reference op by op on a CUDA device
// (every product/sum rounded separately — no FMA contraction — and tensor/scalar division as multiplication by the
// fp32 reciprocal, which is what torch's CUDA div kernel does), so cell positions are bit-identical to the reference's.
// Random numbers are inputs (uniform [0,1) floats / integer coordinates drawn by the caller): the kernels are
// deterministic functions and the oracle can replay them.
#include "common.cuh"
#include "morton.cuh"
#include <algorithm>

namespace ngp {

static constexpr uint32_t DG_TPB = 256;
static constexpr uint32_t DG_CELLS_PER_THREAD = 8;
static constexpr uint32_t DG_CHUNK = DG_TPB * DG_CELLS_PER_THREAD;   // cells per block in the ordered compaction
static constexpr uint32_t DG_INVALID = 0xffffffffu;

struct CascadeScale {
    float s;     // bound_c - half_grid_size   (python double arithmetic, rounded once to fp32 as the tensor*scalar op does)
    float hgs;   // half_grid_size = bound_c / H
    float hgs2;  // 2 * half_grid_size
};
__device__ __forceinline__ CascadeScale cascade_scale(uint32_t cas, float bound, uint32_t H) {
    const double bc = fmin((double)(1u << cas), (double)bound);      // renderer.py:416,472,503  min(2 ** cas, self.bound)
    const double hg = bc / (double)H;
    return {(float)(bc - hg), (float)hg, (float)(hg * 2.0)};
}
// 2 * c / (H - 1) - 1 per axis (renderer.py:413,468,499)
__device__ __forceinline__ float cell_axis(uint32_t c, float inv_hm1) {
    return __fsub_rn(__fmul_rn(__fmul_rn(2.0f, (float)c), inv_hm1), 1.0f);
}
// cas_xyzs = xyzs * (bound - hgs);  cas_xyzs += (u * 2 - 1) * hgs   (renderer.py:474-476, 505-507)
__device__ __forceinline__ float jitter_axis(float a, float u, const CascadeScale& cs, bool has_noise) {
    const float p = __fmul_rn(a, cs.s);
    if (!has_noise) return p;
    return __fadd_rn(p, __fmul_rn(__fsub_rn(__fmul_rn(u, 2.0f), 1.0f), cs.hgs));
}

// ---- mark_untrained_grid -----------------------------------------------------------------------------------------
// One thread per (cascade, Morton cell); the camera poses stream through shared memory in chunks.  A cell is covered by a
// camera when its centre lies in front of it and inside the (slightly dilated) image frustum.
static constexpr uint32_t DG_POSE_CHUNK = 128;
__global__ void __launch_bounds__(DG_TPB) k_dg_mark_untrained(const float* __restrict__ poses, uint32_t B, float rx, float ry,
                                                              float bound, uint32_t C, uint32_t H, uint32_t H3, float inv_hm1,
                                                              float* __restrict__ grid, uint32_t* __restrict__ count_out,
                                                              uint32_t* __restrict__ n_marked) {
    __shared__ float sp[DG_POSE_CHUNK * 12];
    const uint32_t t = blockIdx.x * DG_TPB + threadIdx.x;
    const bool live = t < C * H3;
    const uint32_t cas = live ? t / H3 : 0, i = live ? t % H3 : 0;
    const CascadeScale cs = cascade_scale(cas, bound, H);
    const float wx = __fmul_rn(cell_axis(compact3(i), inv_hm1), cs.s);
    const float wy = __fmul_rn(cell_axis(compact3(i >> 1), inv_hm1), cs.s);
    const float wz = __fmul_rn(cell_axis(compact3(i >> 2), inv_hm1), cs.s);
    uint32_t count = 0;
    for (uint32_t b0 = 0; b0 < B; b0 += DG_POSE_CHUNK) {
        const uint32_t nb = min(DG_POSE_CHUNK, B - b0);
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < nb * 12; k += DG_TPB) {
            const uint32_t b = k / 12, e = k % 12;               // rows 0..2 of the 4x4 c2w matrix: [R | t]
            sp[k] = __ldg(poses + (size_t)(b0 + b) * 16 + (e / 4) * 4 + (e % 4));
        }
        __syncthreads();
        if (!live) continue;
        for (uint32_t b = 0; b < nb; ++b) {
            const float* P = sp + b * 12;                         // P[4*r + c] = pose[r][c], c = 3 is the translation
            const float dx = __fsub_rn(wx, P[3]), dy = __fsub_rn(wy, P[7]), dz = __fsub_rn(wz, P[11]);
            // cam = d @ R  (renderer.py:428-429): cam_j = sum_i d_i R[i][j]
            const float cx_ = fmaf(dz, P[8], fmaf(dy, P[4], __fmul_rn(dx, P[0])));
            const float cy_ = fmaf(dz, P[9], fmaf(dy, P[5], __fmul_rn(dx, P[1])));
            const float cz_ = fmaf(dz, P[10], fmaf(dy, P[6], __fmul_rn(dx, P[2])));
            const bool mz = cz_ > 0.0f;
            const bool mx = fabsf(cx_) < __fadd_rn(__fmul_rn(rx, cz_), cs.hgs2);
            const bool my = fabsf(cy_) < __fadd_rn(__fmul_rn(ry, cz_), cs.hgs2);
            count += (mz && mx && my) ? 1u : 0u;
        }
    }
    if (live) {
        if (count_out) count_out[t] = count;
        if (count == 0) grid[t] = -1.0f;
    }
    if (n_marked) {
        const uint32_t m = __popc(__ballot_sync(0xffffffffu, live && count == 0));
        if ((threadIdx.x & 31u) == 0 && m) atomicAdd(n_marked, m);
    }
}

// ---- ordered list of occupied cells (torch.nonzero(density_grid[cas] > 0)) ----------------------------------------
__device__ __forceinline__ uint32_t occ_flags(const float* __restrict__ g, uint32_t base, uint32_t H3) {
    uint32_t f = 0;
    if (base + DG_CELLS_PER_THREAD <= H3) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(g + base));
        const float4 b = __ldg(reinterpret_cast<const float4*>(g + base) + 1);
        f = (a.x > 0.f ? 1u : 0u) | (a.y > 0.f ? 2u : 0u) | (a.z > 0.f ? 4u : 0u) | (a.w > 0.f ? 8u : 0u) |
            (b.x > 0.f ? 16u : 0u) | (b.y > 0.f ? 32u : 0u) | (b.z > 0.f ? 64u : 0u) | (b.w > 0.f ? 128u : 0u);
    } else {
        for (uint32_t k = 0; k < DG_CELLS_PER_THREAD; ++k)
            if (base + k < H3 && g[base + k] > 0.f) f |= 1u << k;
    }
    return f;
}
// exclusive prefix of `v` over the block (threads in order); returns the block total through *total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[DG_TPB / 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (uint32_t w = 0; w < DG_TPB / 32; ++w) {
        const uint32_t s = warp_sums[w];
        if (w < warp) before += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return before + inc - v;
}
__global__ void __launch_bounds__(DG_TPB) k_dg_occ_count(const float* __restrict__ grid, uint32_t H3, uint32_t nblk,
                                                         uint32_t* __restrict__ block_counts) {
    const uint32_t cas = blockIdx.y;
    const uint32_t base = blockIdx.x * DG_CHUNK + threadIdx.x * DG_CELLS_PER_THREAD;
    const uint32_t c = __popc(occ_flags(grid + (size_t)cas * H3, base, H3));
    uint32_t total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) block_counts[cas * nblk + blockIdx.x] = total;
}
__global__ void __launch_bounds__(DG_TPB) k_dg_occ_scan(uint32_t nblk, uint32_t* __restrict__ block_counts,
                                                        uint32_t* __restrict__ occ_count) {
    const uint32_t cas = blockIdx.x;
    uint32_t* bc = block_counts + cas * nblk;
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < nblk; b0 += DG_TPB) {
        const uint32_t k = b0 + threadIdx.x;
        const uint32_t v = k < nblk ? bc[k] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, &total);
        if (k < nblk) bc[k] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) occ_count[cas] = carry;
}
__global__ void __launch_bounds__(DG_TPB) k_dg_occ_write(const float* __restrict__ grid, uint32_t H3, uint32_t nblk,
                                                         const uint32_t* __restrict__ block_offsets,
                                                         uint32_t* __restrict__ occ_list) {
    const uint32_t cas = blockIdx.y;
    const uint32_t base = blockIdx.x * DG_CHUNK + threadIdx.x * DG_CELLS_PER_THREAD;
    uint32_t f = occ_flags(grid + (size_t)cas * H3, base, H3);
    uint32_t total;
    uint32_t pos = block_offsets[cas * nblk + blockIdx.x] + block_exclusive_scan(__popc(f), &total);
    uint32_t* out = occ_list + (size_t)cas * H3;
    while (f) {
        const uint32_t k = __ffs(f) - 1;
        f &= f - 1;
        out[pos++] = base + k;
    }
}

// ---- sample positions ---------------------------------------------------------------------------------------------
// full update: every cell of every cascade, thread index = Morton index (so the scatter is the identity); the jitter of
// cell (x,y,z) is noise[cas][(x*H + y)*H + z] — the order in which the reference's meshgrid enumerates cells, so the same
// torch.rand stream reproduces the reference's samples.
__global__ void __launch_bounds__(DG_TPB) k_dg_sample_full(uint32_t C, uint32_t H, uint32_t H3, float bound, float inv_hm1,
                                                           const float* __restrict__ noise, float* __restrict__ xyzs) {
    const uint32_t t = blockIdx.x * DG_TPB + threadIdx.x;
    if (t >= C * H3) return;
    const uint32_t cas = t / H3, i = t % H3;
    const CascadeScale cs = cascade_scale(cas, bound, H);
    const uint32_t x = compact3(i), y = compact3(i >> 1), z = compact3(i >> 2);
    float u0 = 0.f, u1 = 0.f, u2 = 0.f;
    if (noise) {
        const float* nz = noise + ((size_t)cas * H3 + ((size_t)x * H + y) * H + z) * 3;
        u0 = __ldg(nz); u1 = __ldg(nz + 1); u2 = __ldg(nz + 2);
    }
    float* o = xyzs + (size_t)t * 3;
    o[0] = jitter_axis(cell_axis(x, inv_hm1), u0, cs, noise != nullptr);
    o[1] = jitter_axis(cell_axis(y, inv_hm1), u1, cs, noise != nullptr);
    o[2] = jitter_axis(cell_axis(z, inv_hm1), u2, cs, noise != nullptr);
}
// partial update: per cascade N uniformly drawn cells followed by N cells drawn from the occupied list.
__global__ void __launch_bounds__(DG_TPB) k_dg_sample_partial(uint32_t C, uint32_t H, uint32_t H3, float bound, float inv_hm1,
                                                              uint32_t N, const int32_t* __restrict__ coords_rand,
                                                              const int64_t* __restrict__ occ_pick_idx,
                                                              const float* __restrict__ occ_pick_u,
                                                              const uint32_t* __restrict__ occ_list,
                                                              const uint32_t* __restrict__ occ_count,
                                                              const float* __restrict__ noise, float* __restrict__ xyzs,
                                                              uint32_t* __restrict__ indices) {
    const uint32_t t = blockIdx.x * DG_TPB + threadIdx.x;
    if (t >= C * 2 * N) return;
    const uint32_t cas = t / (2 * N), n = t % (2 * N);
    const CascadeScale cs = cascade_scale(cas, bound, H);
    uint32_t x, y, z, idx;
    if (n < N) {
        const int32_t* c = coords_rand + ((size_t)cas * N + n) * 3;
        x = (uint32_t)__ldg(c); y = (uint32_t)__ldg(c + 1); z = (uint32_t)__ldg(c + 2);
        idx = (x < H && y < H && z < H) ? morton_enc(x, y, z) : DG_INVALID;
    } else {
        const uint32_t m = n - N, nz = __ldg(occ_count + cas);
        idx = DG_INVALID;
        if (nz > 0) {
            uint32_t pick;
            if (occ_pick_idx) {
                const int64_t p = __ldg(occ_pick_idx + (size_t)cas * N + m);
                pick = (p >= 0 && p < (int64_t)nz) ? (uint32_t)p : DG_INVALID;
            } else {
                pick = min((uint32_t)(__ldg(occ_pick_u + (size_t)cas * N + m) * (float)nz), nz - 1);
            }
            if (pick != DG_INVALID) idx = __ldg(occ_list + (size_t)cas * H3 + pick);
        }
        x = compact3(idx); y = compact3(idx >> 1); z = compact3(idx >> 2);
    }
    float* o = xyzs + (size_t)t * 3;
    indices[t] = idx;
    if (idx == DG_INVALID) { o[0] = o[1] = o[2] = 0.f; return; }
    float u0 = 0.f, u1 = 0.f, u2 = 0.f;
    if (noise) { const float* nzp = noise + (size_t)t * 3; u0 = __ldg(nzp); u1 = __ldg(nzp + 1); u2 = __ldg(nzp + 2); }
    o[0] = jitter_axis(cell_axis(x, inv_hm1), u0, cs, noise != nullptr);
    o[1] = jitter_axis(cell_axis(y, inv_hm1), u1, cs, noise != nullptr);
    o[2] = jitter_axis(cell_axis(z, inv_hm1), u2, cs, noise != nullptr);
}

// ---- update: scatter -> EMA-max + mean -> threshold -> bitfield ---------------------------------------------------
__global__ void k_dg_fill(float4* __restrict__ p, size_t n4, float v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_float4(v, v, v, v);
}
// tmp_grid[cas, idx] = sigma * density_scale (renderer.py:480-482, 511-515).  Cells drawn more than once keep the
// largest of their samples (torch's index_put keeps an arbitrary one): non-negative floats order like their bit
// patterns, the -1 fill is a negative integer, so a signed integer max implements it without a lock.
__global__ void __launch_bounds__(DG_TPB) k_dg_scatter(float* __restrict__ tmp, const uint32_t* __restrict__ indices,
                                                       const float* __restrict__ sigmas, uint32_t C, uint32_t N, uint32_t H3,
                                                       float density_scale) {
    const uint32_t t = blockIdx.x * DG_TPB + threadIdx.x;
    if (t >= C * N) return;
    const uint32_t cas = t / N;
    const uint32_t idx = indices ? __ldg(indices + t) : t % N;
    if (idx >= H3) return;
    const float v = __fadd_rn(__fmul_rn(__ldg(sigmas + t), density_scale), 0.0f);   // + 0 folds -0 into +0
    if (!(v >= 0.0f)) return;                                                         // negative / NaN: tmp stays invalid
    atomicMax(reinterpret_cast<int*>(tmp + (size_t)cas * H3 + idx), __float_as_int(v));
}
// valid = (grid >= 0) & (tmp >= 0); grid[valid] = max(grid * decay, tmp)  (renderer.py:521-523) and the per-block partial
// sums of clamp(grid, 0) for the mean (renderer.py:524), accumulated in double in a fixed order (deterministic).
__device__ __forceinline__ float ema1(float g, float t, float decay) {
    return (g >= 0.0f && t >= 0.0f) ? fmaxf(__fmul_rn(g, decay), t) : g;
}
__global__ void __launch_bounds__(DG_TPB) k_dg_ema(float4* __restrict__ grid, const float4* __restrict__ tmp, size_t n4,
                                                   float decay, double* __restrict__ partials) {
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * DG_TPB + threadIdx.x; i < n4; i += (size_t)gridDim.x * DG_TPB) {
        float4 g = grid[i];
        const float4 t = tmp[i];
        g.x = ema1(g.x, t.x, decay); g.y = ema1(g.y, t.y, decay); g.z = ema1(g.z, t.z, decay); g.w = ema1(g.w, t.w, decay);
        grid[i] = g;
        acc += (double)fmaxf(g.x, 0.f) + (double)fmaxf(g.y, 0.f) + (double)fmaxf(g.z, 0.f) + (double)fmaxf(g.w, 0.f);
    }
    __shared__ double ws[DG_TPB / 32];
#pragma unroll
    for (uint32_t d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31u) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (uint32_t w = 0; w < DG_TPB / 32; ++w) s += ws[w];
        partials[blockIdx.x] = s;
    }
}
// state[0] = mean_density, state[1] = min(mean_density, density_thresh)  (renderer.py:524,529)
__global__ void __launch_bounds__(DG_TPB) k_dg_finalize(const double* __restrict__ partials, uint32_t nparts, double n,
                                                        float density_thresh, float* __restrict__ state) {
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < nparts; i += DG_TPB) acc += partials[i];
    __shared__ double ws[DG_TPB / 32];
#pragma unroll
    for (uint32_t d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31u) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (uint32_t w = 0; w < DG_TPB / 32; ++w) s += ws[w];
        const float mean = (float)(s / n);
        state[0] = mean;
        state[1] = fminf(mean, density_thresh);
    }
}
// packbits with the threshold read from device memory (raymarching.cu:267-289 semantics: bit k of byte n = grid[8n+k] > thresh)
__global__ void __launch_bounds__(DG_TPB) k_dg_packbits(const float4* __restrict__ grid, uint32_t nbytes,
                                                        const float* __restrict__ thresh_p, uint8_t* __restrict__ bitfield) {
    const uint32_t n = blockIdx.x * DG_TPB + threadIdx.x;
    if (n >= nbytes) return;
    const float thresh = __ldg(thresh_p);
    const float4 a = grid[(size_t)n * 2], b = grid[(size_t)n * 2 + 1];
    const uint32_t bits = (a.x > thresh ? 1u : 0u) | (a.y > thresh ? 2u : 0u) | (a.z > thresh ? 4u : 0u) |
                          (a.w > thresh ? 8u : 0u) | (b.x > thresh ? 16u : 0u) | (b.y > thresh ? 32u : 0u) |
                          (b.z > thresh ? 64u : 0u) | (b.w > thresh ? 128u : 0u);
    bitfield[n] = (uint8_t)bits;
}

static int check_grid_dims(const char* who, uint32_t C, uint32_t H) {
    if (C < 1 || C > 24) return fail(NGP_EINVAL, "%s: cascade count out of range", who);
    if (H < 2 || H > 1024 || (H & (H - 1))) return fail(NGP_EINVAL, "%s: grid size must be a power of two in [2, 1024]", who);
    if ((uint64_t)C * H * H * H > 0x7fffffffull) return fail(NGP_EINVAL, "%s: grid too large", who);
    return NGP_OK;
}
static uint32_t ema_blocks(size_t n4) { return (uint32_t)std::min<size_t>(div_up<size_t>(n4, DG_TPB), (size_t)sm_count() * 8); }

}  // namespace ngp

using namespace ngp;

extern "C" int ngp_density_grid_mark_untrained(const float* poses, uint32_t B, float fx, float fy, float cx, float cy,
                                               float bound, uint32_t C, uint32_t H, float* density_grid,
                                               uint32_t* count_out, uint32_t* n_marked, ngp_stream_t stream) {
    if (int rc = check_grid_dims("density_grid_mark_untrained", C, H)) return rc;
    const uint32_t H3 = H * H * H;
    const float rx = (float)((double)cx / (double)fx), ry = (float)((double)cy / (double)fy);   // python: cx / fx (double)
    k_dg_mark_untrained<<<div_up(C * H3, DG_TPB), DG_TPB, 0, as_stream(stream)>>>(poses, B, rx, ry, bound, C, H, H3,
                                                                                  1.0f / (float)(H - 1), density_grid,
                                                                                  count_out, n_marked);
    return check_launch("density_grid_mark_untrained");
}

extern "C" size_t ngp_density_grid_occupied_scratch_bytes(uint32_t C, uint32_t H) {
    return (size_t)C * div_up(H * H * H, DG_CHUNK) * sizeof(uint32_t);
}
extern "C" int ngp_density_grid_occupied(const float* density_grid, uint32_t C, uint32_t H, uint32_t* occ_list,
                                         uint32_t* occ_count, void* scratch, ngp_stream_t stream) {
    if (int rc = check_grid_dims("density_grid_occupied", C, H)) return rc;
    if (H < 4) return fail(NGP_EINVAL, "density_grid_occupied: grid size must be at least 4");
    const uint32_t H3 = H * H * H, nblk = div_up(H3, DG_CHUNK);
    uint32_t* bc = static_cast<uint32_t*>(scratch);
    k_dg_occ_count<<<dim3(nblk, C), DG_TPB, 0, as_stream(stream)>>>(density_grid, H3, nblk, bc);
    if (int rc = check_launch("density_grid_occupied(count)")) return rc;
    k_dg_occ_scan<<<C, DG_TPB, 0, as_stream(stream)>>>(nblk, bc, occ_count);
    if (int rc = check_launch("density_grid_occupied(scan)")) return rc;
    k_dg_occ_write<<<dim3(nblk, C), DG_TPB, 0, as_stream(stream)>>>(density_grid, H3, nblk, bc, occ_list);
    return check_launch("density_grid_occupied(write)");
}

extern "C" int ngp_density_grid_sample_full(uint32_t C, uint32_t H, float bound, const float* noise, float* xyzs,
                                            ngp_stream_t stream) {
    if (int rc = check_grid_dims("density_grid_sample_full", C, H)) return rc;
    const uint32_t H3 = H * H * H;
    k_dg_sample_full<<<div_up(C * H3, DG_TPB), DG_TPB, 0, as_stream(stream)>>>(C, H, H3, bound, 1.0f / (float)(H - 1), noise, xyzs);
    return check_launch("density_grid_sample_full");
}
extern "C" int ngp_density_grid_sample_partial(uint32_t C, uint32_t H, float bound, uint32_t N, const int32_t* coords_rand,
                                               const int64_t* occ_pick_idx, const float* occ_pick_u,
                                               const uint32_t* occ_list, const uint32_t* occ_count, const float* noise,
                                               float* xyzs, uint32_t* indices, ngp_stream_t stream) {
    if (int rc = check_grid_dims("density_grid_sample_partial", C, H)) return rc;
    if (!occ_pick_idx && !occ_pick_u) return fail(NGP_EINVAL, "density_grid_sample_partial: need occ_pick_idx or occ_pick_u");
    if (N == 0) return NGP_OK;
    if ((uint64_t)C * 2 * N > 0x7fffffffull) return fail(NGP_EINVAL, "density_grid_sample_partial: too many samples");
    k_dg_sample_partial<<<div_up(C * 2 * N, DG_TPB), DG_TPB, 0, as_stream(stream)>>>(
        C, H, H * H * H, bound, 1.0f / (float)(H - 1), N, coords_rand, occ_pick_idx, occ_pick_u, occ_list, occ_count, noise,
        xyzs, indices);
    return check_launch("density_grid_sample_partial");
}

extern "C" size_t ngp_density_grid_update_scratch_bytes(uint32_t C, uint32_t H) {
    const size_t n4 = (size_t)C * H * H * H / 4;
    return (size_t)ema_blocks(n4) * sizeof(double);
}
extern "C" int ngp_density_grid_update(float* density_grid, float* tmp_grid, const uint32_t* indices, const float* sigmas,
                                       uint32_t N, float density_scale, float decay, float density_thresh, uint32_t C,
                                       uint32_t H, uint8_t* bitfield, float* state, void* scratch, ngp_stream_t stream) {
    if (int rc = check_grid_dims("density_grid_update", C, H)) return rc;
    const uint32_t H3 = H * H * H;
    if (H3 % 8) return fail(NGP_EINVAL, "density_grid_update: H^3 must be a multiple of 8");
    if (!indices && N != H3) return fail(NGP_EINVAL, "density_grid_update: indices may be NULL only for a full update (N == H^3)");
    if ((uint64_t)C * N > 0x7fffffffull) return fail(NGP_EINVAL, "density_grid_update: too many samples");
    const size_t n = (size_t)C * H3, n4 = n / 4;
    cudaStream_t st = as_stream(stream);
    const uint32_t nb = ema_blocks(n4);
    k_dg_fill<<<nb, DG_TPB, 0, st>>>(reinterpret_cast<float4*>(tmp_grid), n4, -1.0f);
    if (int rc = check_launch("density_grid_update(fill)")) return rc;
    if (N > 0) {
        k_dg_scatter<<<div_up(C * N, DG_TPB), DG_TPB, 0, st>>>(tmp_grid, indices, sigmas, C, N, H3, density_scale);
        if (int rc = check_launch("density_grid_update(scatter)")) return rc;
    }
    double* partials = static_cast<double*>(scratch);
    k_dg_ema<<<nb, DG_TPB, 0, st>>>(reinterpret_cast<float4*>(density_grid), reinterpret_cast<const float4*>(tmp_grid), n4, decay,
                                    partials);
    if (int rc = check_launch("density_grid_update(ema)")) return rc;
    k_dg_finalize<<<1, DG_TPB, 0, st>>>(partials, nb, (double)n, density_thresh, state);
    if (int rc = check_launch("density_grid_update(finalize)")) return rc;
    k_dg_packbits<<<div_up((uint32_t)(n / 8), DG_TPB), DG_TPB, 0, st>>>(reinterpret_cast<const float4*>(density_grid), (uint32_t)(n / 8),
                                                                        state + 1, bitfield);
    return check_launch("density_grid_update(packbits)");
}
