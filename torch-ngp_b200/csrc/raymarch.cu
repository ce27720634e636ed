// raymarch.cu — occupancy-grid ray marcher + volumetric compositor for sm_100a.
//
// Replaces raymarching/src/raymarching.cu: near_far_from_aabb (:91-145), sph_from_ray (:162-198),
// morton3D / morton3D_invert (:214-254), packbits (:267-289), march_rays_train (:311-480),
// composite_rays_train fwd/bwd (:500-682), march_rays (:700-805), composite_rays (:818-905).
//
// Bit-exactness contract (north_star: ray ids / sample counts bit-exact): a ray's march is a
// sequential float recurrence, so the per-step arithmetic is kept operation-for-operation equal to
// what nvcc emits for the reference — every multiply-add the reference's build contracts into an
// FFMA is written here as an explicit fmaf(), divisions are IEEE, the cell index goes through the
// same float->double->float->int chain (:374-376).  What changes is the machinery around it:
//   * slot allocation uses one warp-aggregated atomic pair per warp (shuffle prefix scan + ballot)
//     instead of two atomics per ray; rays of a warp therefore land in ray order (a valid instance
//     of the reference's arbitrary atomics order),
//   * the occupancy bitfield is read through the read-only L1 path,
//   * compositing is unchanged per ray (sequential by definition) but loads are vectorised.
#include "common.cuh"
#include <float.h>

#include "morton.cuh"

namespace ngp {

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
__device__ __forceinline__ float signf1(float x) { return copysignf(1.0f, x); }

// cascade level from position / step size (raymarching.cu:42-54)
__device__ __forceinline__ int mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0.f, (float)e));
}
__device__ __forceinline__ int mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = dt * H * 0.5f;   // reference: (dt*H) in float, *0.5 in double (exact)
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0.f, (float)e));
}

// ---- per-ray marching state -----------------------------------------------------------------
struct RayConst {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max, dt_const, rH, H3, fH, fC, Hm1;
    uint32_t H;
};

__device__ __forceinline__ void ray_setup(RayConst& r, const float* __restrict__ o, const float* __restrict__ d,
                                          float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
    r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
    r.rdx = 1.0f / r.dx; r.rdy = 1.0f / r.dy; r.rdz = 1.0f / r.dz;
    r.bound = bound; r.dt_gamma = dt_gamma;
    r.H = H; r.fH = (float)H; r.fC = (float)C; r.Hm1 = (float)(H - 1);
    r.rH = 1.0f / (float)H;
    r.H3 = (float)(H * H * H);
    const float two_sqrt3 = 2 * 1.7320508075688772f;
    r.dt_min = two_sqrt3 / (float)max_steps;
    r.dt_max = two_sqrt3 * (float)(1 << (C - 1)) / (float)H;
    // the step when dt_gamma == 0: clamp(0, dt_min, dt_max) = fminf(dt_max, fmaxf(dt_min, 0)) — dt_max when dt_min > dt_max
    // (max_steps < H / 2^(C-1)), exactly as the reference's clamp evaluates it
    r.dt_const = clampf(0.0f, r.dt_min, r.dt_max);
}

// One marching decision at parameter t: returns true if the cell is occupied (a sample is taken
// at (x,y,z) with step dt); otherwise advances t past the empty cell.  Mirrors :357-399.
// Strength reductions that keep every result bit-identical (the kernel is ALU-bound: ~130 M of these per step):
//   * the reference's float->double->float cell index chain, (float)(0.5 * (double)a * (double)H), is one correctly
//     rounded product of a with the exactly representable constant 0.5*H, i.e. the single FMUL a * (0.5f*H);
//   * with a single cascade (C == 1) both mip_from_* clamp to level 0: frexp/scalbn/reciprocal drop out;
//   * with dt_gamma == 0 the step clamp(t*0, dt_min, dt_max) is the constant dt_const = min(dt_max, max(dt_min, 0)) (also for
//     t = inf: NaN clamps to it through fmaxf/fminf), so the empty-space loop is add/compare only.
__device__ __forceinline__ bool march_probe(const RayConst& r, const uint8_t* __restrict__ grid, float& t,
                                            float& x, float& y, float& z, float& dt) {
    x = clampf(fmaf(t, r.dx, r.ox), -r.bound, r.bound);
    y = clampf(fmaf(t, r.dy, r.oy), -r.bound, r.bound);
    z = clampf(fmaf(t, r.dz, r.oz), -r.bound, r.bound);
    const bool const_dt = (r.dt_gamma == 0.0f);
    dt = const_dt ? r.dt_const : clampf(t * r.dt_gamma, r.dt_min, r.dt_max);

    int level = 0;
    float mip_bound = fminf(1.0f, r.bound), mip_rbound;
    if (r.fC > 1.0f) {
        level = max(mip_from_pos(x, y, z, r.fC), mip_from_dt(dt, r.fH, r.fC));
        mip_bound = fminf(scalbnf(1.0f, level), r.bound);
    }
    mip_rbound = 1.0f / mip_bound;

    const float halfH = 0.5f * r.fH;
    const int nx = (int)clampf(fmaf(x, mip_rbound, 1.0f) * halfH, 0.0f, r.Hm1);
    const int ny = (int)clampf(fmaf(y, mip_rbound, 1.0f) * halfH, 0.0f, r.Hm1);
    const int nz = (int)clampf(fmaf(z, mip_rbound, 1.0f) * halfH, 0.0f, r.Hm1);

    const uint32_t index = (uint32_t)fmaf((float)level, r.H3, (float)morton_enc(nx, ny, nz));
    const bool occ = __ldg(grid + (index >> 3)) & (1u << (index & 7u));
    if (occ) return true;

    const float tx = (fmaf(fmaf(fmaf(0.5f, signf1(r.dx), (float)nx + 0.5f) * r.rH, 2.0f, -1.0f), mip_bound, -x)) * r.rdx;
    const float ty = (fmaf(fmaf(fmaf(0.5f, signf1(r.dy), (float)ny + 0.5f) * r.rH, 2.0f, -1.0f), mip_bound, -y)) * r.rdy;
    const float tz = (fmaf(fmaf(fmaf(0.5f, signf1(r.dz), (float)nz + 0.5f) * r.rH, 2.0f, -1.0f), mip_bound, -z)) * r.rdz;
    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    if (const_dt) {
        do { t += r.dt_const; } while (t < tt);
    } else {
        do { t += clampf(t * r.dt_gamma, r.dt_min, r.dt_max); } while (t < tt);
    }
    return false;
}

// ---- utils ----------------------------------------------------------------------------------
__global__ void k_near_far(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                           const float* __restrict__ aabb, uint32_t N, float min_near,
                           float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float rdx = 1.0f / rays_d[n * 3], rdy = 1.0f / rays_d[n * 3 + 1], rdz = 1.0f / rays_d[n * 3 + 2];

    float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx;
    if (near > far) { const float c = near; near = far; far = c; }
    float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
    if (near_y > far_y) { const float c = near_y; near_y = far_y; far_y = c; }
    if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
    if (near_z > far_z) { const float c = near_z; near_z = far_z; far_z = c; }
    if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    nears[n] = near;
    fars[n] = far;
}

__global__ void k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float radius,
                               uint32_t N, float* __restrict__ coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float Bh = ox * dx + oy * dy + oz * dz;
    const float Cc = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-Bh + sqrtf(Bh * Bh - A * Cc)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    const float rpi = 0.3183098861837907f;
    coords[n * 2] = 2 * theta * rpi - 1;
    coords[n * 2 + 1] = phi * rpi;
}

__global__ void k_morton3D(const int* __restrict__ coords, uint32_t N, int* __restrict__ indices) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    indices[n] = (int)morton_enc(coords[n * 3], coords[n * 3 + 1], coords[n * 3 + 2]);
}

__global__ void k_morton3D_invert(const int* __restrict__ indices, uint32_t N, int* __restrict__ coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const int ind = indices[n];
    coords[n * 3] = (int)compact3((uint32_t)(ind >> 0));
    coords[n * 3 + 1] = (int)compact3((uint32_t)(ind >> 1));
    coords[n * 3 + 2] = (int)compact3((uint32_t)(ind >> 2));
}

// 8 floats -> 1 occupancy byte; two 128-bit loads per thread
__global__ void k_packbits(const float* __restrict__ grid, uint32_t N, float thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float4 a = __ldg(reinterpret_cast<const float4*>(grid) + (size_t)n * 2);
    const float4 b = __ldg(reinterpret_cast<const float4*>(grid) + (size_t)n * 2 + 1);
    uint32_t bits = 0;
    bits |= (a.x > thresh) ? 1u : 0u;   bits |= (a.y > thresh) ? 2u : 0u;
    bits |= (a.z > thresh) ? 4u : 0u;   bits |= (a.w > thresh) ? 8u : 0u;
    bits |= (b.x > thresh) ? 16u : 0u;  bits |= (b.y > thresh) ? 32u : 0u;
    bits |= (b.z > thresh) ? 64u : 0u;  bits |= (b.w > thresh) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

// ---- training marcher ------------------------------------------------------------------------
// The reference marches every ray twice (count, then emit).  Here the first pass records the parameter t of each
// sample it finds (up to MARCH_K per ray) in shared memory; after the warp-aggregated slot reservation the samples
// of the whole warp are emitted COOPERATIVELY: lane j takes the warp's j-th sample (owner ray found by a shuffle
// binary search over the warp's inclusive scan), recomputes position / step size from the cached t with exactly the
// arithmetic of the marching loop, and writes it — consecutive lanes write consecutive samples (coalesced), no
// divergent re-march.  Rays with more than MARCH_K samples fall back to the per-lane second pass.
static constexpr uint32_t MARCH_K = 64;
static constexpr uint32_t MARCH_TPB = 64;              // small blocks: the per-ray work is very uneven (sky vs object), 64-thread
                                                       // CTAs retire sooner and rebalance across SMs
static constexpr uint32_t MARCH_PITCH = MARCH_K + 1;   // odd pitch: conflict-free reads of one ray's consecutive samples

__global__ void __launch_bounds__(MARCH_TPB)
k_march_rays_train(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                   const uint8_t* __restrict__ grid, float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                   uint32_t C, uint32_t H, uint32_t M, const float* __restrict__ nears,
                   const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
                   float* __restrict__ deltas, int* __restrict__ rays, int* __restrict__ counter,
                   const float* __restrict__ noises) {
    __shared__ float tcache[MARCH_TPB * MARCH_PITCH];
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    const bool active = n < N;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wbase = threadIdx.x & ~31u;
    float* my_cache = tcache + threadIdx.x * MARCH_PITCH;

    RayConst r;
    r.ox = r.oy = r.oz = 0.f; r.dx = r.dy = r.dz = 1.f;
    float far = 0.f, t0 = 0.f;
    uint32_t num_steps = 0;
    if (active) {
        ray_setup(r, rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, bound, dt_gamma, max_steps, C, H);
        const float near = nears[n];
        far = fars[n];
        t0 = fmaf(clampf(near * dt_gamma, r.dt_min, r.dt_max), noises[n], near);
        // pass 1: count occupied steps, remember where they were
        float t = t0, x, y, z, dt;
        while (t < far && num_steps < max_steps) {
            if (march_probe(r, grid, t, x, y, z, dt)) {
                if (num_steps < MARCH_K) my_cache[num_steps] = t;
                num_steps++;
                t += dt;
            }
        }
    } else {
        // shared (warp-uniform) marching constants are needed by every lane in the cooperative phase
        const float two_sqrt3 = 2 * 1.7320508075688772f;
        r.bound = bound; r.dt_gamma = dt_gamma;
        r.dt_min = two_sqrt3 / (float)max_steps;
        r.dt_max = two_sqrt3 * (float)(1 << (C - 1)) / (float)H;
    }

    // warp-aggregated slot reservation: inclusive scan of num_steps, ballot-rank of active lanes, one atomic pair
    uint32_t incl = num_steps;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    const uint32_t warp_total = __shfl_sync(FULL, incl, 31);
    const uint32_t act_mask = __ballot_sync(FULL, active);
    if (act_mask == 0) return;
    uint32_t base_pt = 0, base_ray = 0;
    if (lane == 0) {
        base_pt = (uint32_t)atomicAdd(counter, (int)warp_total);
        base_ray = (uint32_t)atomicAdd(counter + 1, (int)__popc(act_mask));
    }
    base_pt = __shfl_sync(FULL, base_pt, 0);
    base_ray = __shfl_sync(FULL, base_ray, 0);

    const uint32_t point_index = base_pt + incl - num_steps;
    if (active) {
        const uint32_t ray_index = base_ray + __popc(act_mask & ((1u << lane) - 1u));
        rays[ray_index * 3] = (int)n;
        rays[ray_index * 3 + 1] = (int)point_index;
        rays[ray_index * 3 + 2] = (int)num_steps;
    }
    const bool fits = active && num_steps > 0 && point_index + num_steps <= M;

    // ---- cooperative emission of the cached rays ----
    __syncwarp();
    const uint32_t cnt = (fits && num_steps <= MARCH_K) ? num_steps : 0u;
    uint32_t cincl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, cincl, o);
        if (lane >= (uint32_t)o) cincl += v;
    }
    const uint32_t ctotal = __shfl_sync(FULL, cincl, 31);
    for (uint32_t j0 = 0; j0 < ctotal; j0 += 32) {
        const uint32_t j = j0 + lane;
        // owner = first lane whose inclusive count exceeds j
        uint32_t o = 0;
#pragma unroll
        for (uint32_t step = 16; step >= 1; step >>= 1) {
            const uint32_t v = __shfl_sync(FULL, cincl, o + step - 1);
            if (v <= j) o += step;
        }
        o &= 31u;
        const uint32_t o_incl = __shfl_sync(FULL, cincl, o), o_cnt = __shfl_sync(FULL, cnt, o);
        const float ox = __shfl_sync(FULL, r.ox, o), oy = __shfl_sync(FULL, r.oy, o), oz = __shfl_sync(FULL, r.oz, o);
        const float dx = __shfl_sync(FULL, r.dx, o), dy = __shfl_sync(FULL, r.dy, o), dz = __shfl_sync(FULL, r.dz, o);
        const float ot0 = __shfl_sync(FULL, t0, o);
        const uint32_t opt = __shfl_sync(FULL, point_index, o);
        if (j < ctotal) {
            const uint32_t sidx = j - (o_incl - o_cnt);
            const float* oc = tcache + (wbase + o) * MARCH_PITCH;
            const float t = oc[sidx];
            float last_t = ot0;
            if (sidx > 0) { const float tp = oc[sidx - 1]; last_t = tp + clampf(tp * dt_gamma, r.dt_min, r.dt_max); }
            const float x = clampf(fmaf(t, dx, ox), -bound, bound);
            const float y = clampf(fmaf(t, dy, oy), -bound, bound);
            const float z = clampf(fmaf(t, dz, oz), -bound, bound);
            const float dt = clampf(t * dt_gamma, r.dt_min, r.dt_max);
            const float t_after = t + dt;
            const size_t out = (size_t)opt + sidx;
            xyzs[out * 3] = x; xyzs[out * 3 + 1] = y; xyzs[out * 3 + 2] = z;
            dirs[out * 3] = dx; dirs[out * 3 + 1] = dy; dirs[out * 3 + 2] = dz;
            *reinterpret_cast<float2*>(deltas + out * 2) = make_float2(dt, t_after - last_t);
        }
    }

    // ---- fallback: rays with more samples than the cache holds re-march on their own ----
    if (fits && num_steps > MARCH_K) {
        float* __restrict__ px = xyzs + (size_t)point_index * 3;
        float* __restrict__ pd = dirs + (size_t)point_index * 3;
        float* __restrict__ pl = deltas + (size_t)point_index * 2;
        float t = t0, last_t = t0, x, y, z, dt;
        uint32_t step = 0;
        while (t < far && step < num_steps) {
            if (march_probe(r, grid, t, x, y, z, dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                t += dt;
                *reinterpret_cast<float2*>(pl) = make_float2(dt, t - last_t);
                last_t = t;
                px += 3; pd += 3; pl += 2;
                step++;
            }
        }
    }
}

// ---- training compositor ---------------------------------------------------------------------
// The reference walks each ray's samples sequentially in one thread (one dependent load + exp per step; a 200-sample
// ray is a ~200-deep latency chain, and neighbouring threads read segments that are far apart).  Here ONE WARP OWNS ONE
// RAY (empty rays retire immediately; a first version gave each warp 32 consecutive rays, which serialised the
// clustered non-empty rays of an image region in the same warp): the 32 lanes load 32 consecutive samples of the ray
// (coalesced), evaluate alpha in parallel, and obtain the running transmittance / colour / depth with warp prefix
// scans (product scan for T, sum scans for t and the accumulated colour).  Early termination (T < T_thresh after a
// sample, that sample included) is found with a ballot.  Results equal the sequential recurrence up to fp32
// re-association inside the scans (tests: 1e-5 of the output scale).
__device__ __forceinline__ float warp_incl_prod(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (uint32_t)o) v *= t; }
    return v;
}
__device__ __forceinline__ float warp_incl_sum(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (uint32_t)o) v += t; }
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct MseArgs {             // optional loss head of k_composite_train_fwd (target == nullptr: plain compositor)
    const float* target;     // [N,3] indexed like image
    float bg, inv_norm;      // background colour; 2 / (3 R)
    const float* scale;      // device scalar: loss scale
    float* g_image;          // [N,3] d(scaled loss)/d image
    float* g_ws;             // [N]   d(scaled loss)/d weights_sum
    float* sqerr;            // [N]   squared error per row of the rays table
};

// G = lanes per ray (8, 16 or 32).  The bench scene has 8 samples per ray on average (most rays that hit anything carry 10-30, half of the
// rays none): with one warp per ray three quarters of the lanes idle and the kernels are latency-bound (ncu r2c: long-scoreboard 61-72 %,
// issue 38-69 %).  A group of G lanes owns a ray and walks it G samples at a time; the 32 / G groups of a warp iterate together until the
// longest of their rays is done (rays of a warp are neighbours in the rays table, i.e. similar lengths).  Same arithmetic per sample, the scans are
// G wide (different re-association than the 32-wide form, same tolerance).
template <int G> __device__ __forceinline__ float grp_incl_prod(float v, uint32_t gl) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, v, o, G); if (gl >= (uint32_t)o) v *= t; }
    return v;
}
template <int G> __device__ __forceinline__ float grp_incl_sum(float v, uint32_t gl) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, v, o, G); if (gl >= (uint32_t)o) v += t; }
    return v;
}
template <int G> __device__ __forceinline__ float grp_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int G>
__global__ void __launch_bounds__(128)
k_composite_train_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                      const float* __restrict__ deltas, const int* __restrict__ rays, uint32_t M, uint32_t N,
                      float T_thresh, float* __restrict__ weights_sum, float* __restrict__ depth,
                      float* __restrict__ image, const MseArgs mse) {
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr uint32_t GMASK = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gl = lane & (G - 1u), gbase = lane - gl;
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) / G;        // one group per row of the rays table
    const bool have_ray = n < N;
    uint32_t index = 0, offset = 0, cnt = 0;
    if (have_ray) { index = rays[n * 3]; offset = rays[n * 3 + 1]; cnt = rays[n * 3 + 2]; }
    const bool ok = have_ray && cnt != 0 && offset + cnt <= M;            // empty / dropped ray: outputs stay 0
    const float* __restrict__ sg = sigmas + offset;
    const float* __restrict__ cl = rgbs + (size_t)offset * 3;
    const float2* __restrict__ dl = reinterpret_cast<const float2*>(deltas) + offset;
    float T = 1.0f, t_acc = 0.f;
    float pr = 0, pg = 0, pb = 0, pws = 0, pd = 0;                       // per-lane partial sums
    bool done = !ok;
    for (uint32_t base = 0;; base += G) {
        const bool mine = !done && base < cnt;
        if (!__any_sync(FULL, mine)) break;                              // warp-uniform: every group of the warp is finished
        const uint32_t i = base + gl;
        const bool valid = mine && i < cnt;
        float alpha = 0.f, d1 = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (valid) {
            const float2 dd = __ldg(dl + i);
            alpha = 1.0f - __expf(-__ldg(sg + i) * dd.x);
            d1 = dd.y;
            c0 = __ldg(cl + i * 3); c1 = __ldg(cl + i * 3 + 1); c2 = __ldg(cl + i * 3 + 2);
        }
        const float p_incl = grp_incl_prod<G>(1.0f - alpha, gl);
        float p_excl = __shfl_up_sync(FULL, p_incl, 1, G);
        if (gl == 0) p_excl = 1.0f;
        const float t_i = t_acc + grp_incl_sum<G>(d1, gl);
        const float T_after = T * p_incl;
        const uint32_t term = (__ballot_sync(FULL, valid && (T_after < T_thresh)) >> gbase) & GMASK;
        const uint32_t last = term ? (uint32_t)(__ffs(term) - 1) : (uint32_t)(G - 1);   // last lane of the group that still contributes
        if (valid && gl <= last) {
            const float w = alpha * (T * p_excl);
            pr = fmaf(w, c0, pr); pg = fmaf(w, c1, pg); pb = fmaf(w, c2, pb);
            pws += w;
            pd = fmaf(w, t_i, pd);
        }
        const float p_last = __shfl_sync(FULL, p_incl, G - 1, G), t_last = __shfl_sync(FULL, t_i, G - 1, G);
        if (mine) {
            if (term) done = true;
            else { T = T * p_last; t_acc = t_last; }
        }
    }
    const float r = grp_sum<G>(pr), g = grp_sum<G>(pg), b = grp_sum<G>(pb), ws = grp_sum<G>(pws), d = grp_sum<G>(pd);
    if (gl == 0 && have_ray) {
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
        if (mse.target) {
            // loss head of the training step (nerf/utils.py:861-868 with the renderer's background blend, renderer.py:313):
            // pred = image + (1 - ws) * bg ; loss = mean((pred - target)^2) over 3 R values ; gradients scaled by the loss scale
            const float k = mse.inv_norm * __ldg(mse.scale);
            const float bgw = (1.0f - ws) * mse.bg;
            const float d0 = (r + bgw) - __ldg(mse.target + index * 3), d1 = (g + bgw) - __ldg(mse.target + index * 3 + 1),
                        d2 = (b + bgw) - __ldg(mse.target + index * 3 + 2);
            const float g0 = d0 * k, g1 = d1 * k, g2 = d2 * k;
            mse.g_image[index * 3] = g0; mse.g_image[index * 3 + 1] = g1; mse.g_image[index * 3 + 2] = g2;
            mse.g_ws[index] = -((g0 + g1) + g2) * mse.bg;
            mse.sqerr[n] = (d0 * d0 + d1 * d1) + d2 * d2;
        }
    }
}

// tail of the marcher in the step driver: the reference's 16-slot sample-count ring (renderer.py:281-283), advanced on the device
__global__ void k_step_counter_push(int* __restrict__ ring, const int* __restrict__ counter, int* __restrict__ nsteps,
                                    int* __restrict__ step_counter) {
    const int r = ring[0] & 15;
    step_counter[2 * r] = counter[0]; step_counter[2 * r + 1] = counter[1];
    ring[0] = (r + 1) & 15;
    nsteps[0] += 1;
}

template <int G>
__global__ void __launch_bounds__(128)
k_composite_train_bwd(const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image,
                      const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                      const float* __restrict__ deltas, const int* __restrict__ rays,
                      const float* __restrict__ weights_sum, const float* __restrict__ image, uint32_t M,
                      uint32_t N, float T_thresh, float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr uint32_t GMASK = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gl = lane & (G - 1u), gbase = lane - gl;
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) / G;        // one group per ray
    const bool have_ray = n < N;
    uint32_t index = 0, offset = 0, cnt = 0;
    if (have_ray) { index = rays[n * 3]; offset = rays[n * 3 + 1]; cnt = rays[n * 3 + 2]; }
    const bool ok = have_ray && cnt != 0 && offset + cnt <= M;
    float gws = 0.f, gr = 0.f, gg = 0.f, gb = 0.f, r_final = 0.f, g_final = 0.f, b_final = 0.f, ws_term = 0.f;
    if (ok) {
        gws = __ldg(grad_weights_sum + index);
        gr = __ldg(grad_image + index * 3); gg = __ldg(grad_image + index * 3 + 1); gb = __ldg(grad_image + index * 3 + 2);
        r_final = __ldg(image + index * 3); g_final = __ldg(image + index * 3 + 1); b_final = __ldg(image + index * 3 + 2);
        ws_term = gws * (1 - __ldg(weights_sum + index));
    }
    const float* __restrict__ sg = sigmas + offset;
    const float* __restrict__ cl = rgbs + (size_t)offset * 3;
    const float2* __restrict__ dl = reinterpret_cast<const float2*>(deltas) + offset;
    float* __restrict__ gs = grad_sigmas + offset;
    float* __restrict__ gc = grad_rgbs + (size_t)offset * 3;
    float T = 1.0f, r_acc = 0.f, g_acc = 0.f, b_acc = 0.f;
    bool done = !ok;
    for (uint32_t base = 0;; base += G) {
        const bool mine = !done && base < cnt;
        if (!__any_sync(FULL, mine)) break;
        const uint32_t i = base + gl;
        const bool valid = mine && i < cnt;
        float alpha = 0.f, d0 = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (valid) {
            d0 = __ldg(dl + i).x;
            alpha = 1.0f - __expf(-__ldg(sg + i) * d0);
            c0 = __ldg(cl + i * 3); c1 = __ldg(cl + i * 3 + 1); c2 = __ldg(cl + i * 3 + 2);
        }
        const float p_incl = grp_incl_prod<G>(1.0f - alpha, gl);
        float p_excl = __shfl_up_sync(FULL, p_incl, 1, G);
        if (gl == 0) p_excl = 1.0f;
        const float w = alpha * (T * p_excl);
        const float T_after = T * p_incl;
        // colour accumulated up to and including sample i
        const float r_i = r_acc + grp_incl_sum<G>(w * c0, gl);
        const float g_i = g_acc + grp_incl_sum<G>(w * c1, gl);
        const float b_i = b_acc + grp_incl_sum<G>(w * c2, gl);
        const uint32_t term = (__ballot_sync(FULL, valid && (T_after < T_thresh)) >> gbase) & GMASK;
        const uint32_t last = term ? (uint32_t)(__ffs(term) - 1) : (uint32_t)(G - 1);
        if (valid && gl <= last) {
            gc[i * 3] = gr * w; gc[i * 3 + 1] = gg * w; gc[i * 3 + 2] = gb * w;
            gs[i] = d0 * (gr * (T_after * c0 - (r_final - r_i)) + gg * (T_after * c1 - (g_final - g_i)) +
                          gb * (T_after * c2 - (b_final - b_i)) + ws_term);
        }
        const float p_last = __shfl_sync(FULL, p_incl, G - 1, G);
        const float r_last = __shfl_sync(FULL, r_i, G - 1, G), g_last = __shfl_sync(FULL, g_i, G - 1, G), b_last = __shfl_sync(FULL, b_i, G - 1, G);
        if (mine) {
            if (term) done = true;
            else { T = T * p_last; r_acc = r_last; g_acc = g_last; b_acc = b_last; }
        }
    }
}

// ---- inference -------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_march_rays(uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive, const float* __restrict__ rays_t,
             const float* __restrict__ rays_o, const float* __restrict__ rays_d, float bound, float dt_gamma,
             uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* __restrict__ grid,
             const float* __restrict__ nears, const float* __restrict__ fars, float* __restrict__ xyzs,
             float* __restrict__ dirs, float* __restrict__ deltas, const float* __restrict__ noises) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    RayConst r;
    ray_setup(r, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, bound, dt_gamma, max_steps, C, H);
    float* __restrict__ px = xyzs + (size_t)n * n_step * 3;
    float* __restrict__ pd = dirs + (size_t)n * n_step * 3;
    float* __restrict__ pl = deltas + (size_t)n * n_step * 2;
    float t = rays_t[index];
    const float far = fars[index];
    (void)nears;
    t = fmaf(clampf(t * dt_gamma, r.dt_min, r.dt_max), noises[n], t);
    float last_t = t, x, y, z, dt;
    uint32_t step = 0;
    while (t < far && step < n_step) {
        if (march_probe(r, grid, t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            t += dt;
            pl[0] = dt; pl[1] = t - last_t;
            last_t = t;
            px += 3; pd += 3; pl += 2;
            step++;
        }
    }
}

__global__ void __launch_bounds__(128)
k_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int* __restrict__ rays_alive,
                 float* __restrict__ rays_t, const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                 const float* __restrict__ deltas, float* __restrict__ weights_sum, float* __restrict__ depth,
                 float* __restrict__ image) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const float* __restrict__ sg = sigmas + (size_t)n * n_step;
    const float* __restrict__ cl = rgbs + (size_t)n * n_step * 3;
    const float* __restrict__ dl = deltas + (size_t)n * n_step * 2;
    float t = rays_t[index];
    float weight_sum = weights_sum[index], d = depth[index];
    float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
    uint32_t step = 0;
    while (step < n_step) {
        const float d0 = dl[step * 2];
        if (d0 == 0) break;
        const float alpha = 1.0f - __expf(-sg[step] * d0);
        const float T = 1 - weight_sum;
        const float weight = alpha * T;
        weight_sum += weight;
        t += dl[step * 2 + 1];
        d = fmaf(weight, t, d);
        r = fmaf(weight, cl[step * 3], r);
        g = fmaf(weight, cl[step * 3 + 1], g);
        b = fmaf(weight, cl[step * 3 + 2], b);
        if (T < T_thresh) break;
        step++;
    }
    if (step < n_step) rays_alive[n] = -1;
    else rays_t[index] = t;
    weights_sum[index] = weight_sum;
    depth[index] = d;
    image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
}

// ---- device-driven inference loop (extension) ----------------------------------------------------------------------------
// The reference's eval loop (renderer.py:337-367) reads n_alive back to the host every iteration (`rays_alive[rays_alive >= 0]`, a
// boolean-index copy + synchronisation) to size the next launch.  Here the loop state lives in a device control block
//   ctrl[0] = n_alive   ctrl[1] = n_step = max(min(N / n_alive, 8), 1)   ctrl[2] = M = n_alive * n_step rounded up to 128 (+128 when
//   already a multiple, the wrapper's rule)   ctrl[3] = samples marched per ray so far (loop ends at max_steps)   ctrl[4] = scratch
// and every kernel is launched for the worst case (N rays) and returns early past n_alive, so an iteration needs no host
// round trip and a block of iterations can be captured in one CUDA graph.  Per-ray arithmetic is the kernels' above (same
// march_probe / compositing sequence, same n_step schedule); only the ORDER of the compacted ray list differs (atomics), which no
// per-ray result depends on.
__global__ void k_infer_init(uint32_t N, int* __restrict__ rays_alive, uint32_t* __restrict__ ctrl) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n < N) rays_alive[n] = (int)n;
    if (n == 0) {
        ctrl[0] = N; ctrl[1] = 1u;
        ctrl[2] = N + (128u - N % 128u);
        ctrl[3] = 0u; ctrl[4] = 0u;
    }
}

__global__ void __launch_bounds__(128)
k_march_rays_dev(const uint32_t* __restrict__ ctrl, const int* __restrict__ rays_alive, const float* __restrict__ rays_t,
                 const float* __restrict__ rays_o, const float* __restrict__ rays_d, float bound, float dt_gamma,
                 uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* __restrict__ grid,
                 const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
                 const float* __restrict__ noises) {
    const uint32_t n_alive = ctrl[0], n_step = ctrl[1];
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    RayConst r;
    ray_setup(r, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, bound, dt_gamma, max_steps, C, H);
    float* __restrict__ px = xyzs + (size_t)n * n_step * 3;
    float* __restrict__ pd = dirs + (size_t)n * n_step * 3;
    float* __restrict__ pl = deltas + (size_t)n * n_step * 2;
    float t = rays_t[index];
    const float far = fars[index];
    const float noise = (noises && ctrl[3] == 0u) ? noises[n] : 0.0f;      // perturb only in the first iteration (renderer.py:353)
    t = fmaf(clampf(t * dt_gamma, r.dt_min, r.dt_max), noise, t);
    float last_t = t, x, y, z, dt;
    uint32_t step = 0;
    while (t < far && step < n_step) {
        if (march_probe(r, grid, t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            t += dt;
            pl[0] = dt; pl[1] = t - last_t;
            last_t = t;
            px += 3; pd += 3; pl += 2;
            step++;
        }
    }
    // the reference zero-fills the buffers before every launch (raymarching.py:336-338); only the first unused delta matters (the
    // compositor stops at delta == 0), so write that sentinel instead of clearing M rows
    if (step < n_step) { pl[0] = 0.0f; pl[1] = 0.0f; }
}

__global__ void __launch_bounds__(128)
k_composite_rays_dev(const uint32_t* __restrict__ ctrl, float T_thresh, int* __restrict__ rays_alive, float* __restrict__ rays_t,
                     const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                     float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n_alive = ctrl[0], n_step = ctrl[1];
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const float* __restrict__ sg = sigmas + (size_t)n * n_step;
    const float* __restrict__ cl = rgbs + (size_t)n * n_step * 3;
    const float* __restrict__ dl = deltas + (size_t)n * n_step * 2;
    float t = rays_t[index];
    float weight_sum = weights_sum[index], d = depth[index];
    float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
    uint32_t step = 0;
    while (step < n_step) {
        const float d0 = dl[step * 2];
        if (d0 == 0) break;
        const float alpha = 1.0f - __expf(-sg[step] * d0);
        const float T = 1 - weight_sum;
        const float weight = alpha * T;
        weight_sum += weight;
        t += dl[step * 2 + 1];
        d = fmaf(weight, t, d);
        r = fmaf(weight, cl[step * 3], r);
        g = fmaf(weight, cl[step * 3 + 1], g);
        b = fmaf(weight, cl[step * 3 + 2], b);
        if (T < T_thresh) break;
        step++;
    }
    if (step < n_step) rays_alive[n] = -1;
    else rays_t[index] = t;
    weights_sum[index] = weight_sum;
    depth[index] = d;
    image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
}

// survivors of rays_in[0 .. n_alive) -> rays_out (warp-aggregated slot reservation in ctrl[4])
__global__ void __launch_bounds__(256)
k_compact_rays_dev(uint32_t* __restrict__ ctrl, const int* __restrict__ rays_in, int* __restrict__ rays_out) {
    const uint32_t n_alive = ctrl[0];
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    const int v = (n < n_alive) ? rays_in[n] : -1;
    const bool keep = v >= 0;
    const uint32_t mask = __ballot_sync(0xffffffffu, keep);
    if (mask == 0) return;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(ctrl + 4, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) rays_out[base + __popc(mask & ((1u << lane) - 1u))] = v;
}

// next iteration's control block (one thread): n_alive = survivors, n_step = max(min(N / n_alive, 8), 1) (renderer.py:349)
__global__ void k_infer_advance(uint32_t* __restrict__ ctrl, uint32_t N, uint32_t max_steps) {
    const uint32_t steps_done = ctrl[3] + ctrl[1];
    uint32_t alive = ctrl[4];
    if (steps_done >= max_steps) alive = 0;                 // `while step < max_steps`
    uint32_t n_step = 1;
    if (alive > 0) { n_step = N / alive; n_step = n_step < 8u ? n_step : 8u; n_step = n_step > 1u ? n_step : 1u; }
    const uint32_t m = alive * n_step;
    ctrl[0] = alive; ctrl[1] = n_step;
    ctrl[2] = alive ? m + (128u - m % 128u) : 0u;
    ctrl[3] = steps_done; ctrl[4] = 0u;
}

}  // namespace ngp

using namespace ngp;

#define NGP_LAUNCH_1D(KERNEL, COUNT, TPB, NAME, ...)                                   \
    do {                                                                               \
        if ((COUNT) == 0) return NGP_OK;                                               \
        KERNEL<<<div_up((uint32_t)(COUNT), (uint32_t)(TPB)), (TPB), 0, as_stream(stream)>>>(__VA_ARGS__); \
        return check_launch(NAME);                                                     \
    } while (0)

// training compositor: lanes per ray (NGP_COMPOSITE_GROUP = 8 | 16 | 32, default 8; 32 = one warp per ray, the round-1 form)
static int composite_group() {
    static const int g = [] { const char* e = getenv("NGP_COMPOSITE_GROUP"); const int v = e ? atoi(e) : 8; return (v == 16 || v == 32) ? v : 8; }();
    return g;
}
#define NGP_LAUNCH_COMPOSITE(KERNEL, NAME, ...)                                                                        \
    do {                                                                                                               \
        if (N == 0) return NGP_OK;                                                                                     \
        const int grp = composite_group();                                                                             \
        if ((uint64_t)N * (uint64_t)grp > 0xffffffffull) return fail(NGP_EINVAL, NAME ": too many rays");             \
        const uint32_t blocks = div_up(N * (uint32_t)grp, 128u);                                                       \
        if (grp == 8) KERNEL<8><<<blocks, 128, 0, as_stream(stream)>>>(__VA_ARGS__);                                   \
        else if (grp == 16) KERNEL<16><<<blocks, 128, 0, as_stream(stream)>>>(__VA_ARGS__);                            \
        else KERNEL<32><<<blocks, 128, 0, as_stream(stream)>>>(__VA_ARGS__);                                           \
        return check_launch(NAME);                                                                                     \
    } while (0)

extern "C" int ngp_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                                      float min_near, float* nears, float* fars, ngp_stream_t stream) {
    NGP_LAUNCH_1D(k_near_far, N, 128, "near_far_from_aabb", rays_o, rays_d, aabb, N, min_near, nears, fars);
}
extern "C" int ngp_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                                ngp_stream_t stream) {
    NGP_LAUNCH_1D(k_sph_from_ray, N, 128, "sph_from_ray", rays_o, rays_d, radius, N, coords);
}
extern "C" int ngp_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, ngp_stream_t stream) {
    NGP_LAUNCH_1D(k_morton3D, N, 256, "morton3D", coords, N, indices);
}
extern "C" int ngp_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, ngp_stream_t stream) {
    NGP_LAUNCH_1D(k_morton3D_invert, N, 256, "morton3D_invert", indices, N, coords);
}
extern "C" int ngp_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield,
                            ngp_stream_t stream) {
    if ((reinterpret_cast<uintptr_t>(grid) & 15) != 0) return fail(NGP_EINVAL, "packbits: grid must be 16-byte aligned");
    NGP_LAUNCH_1D(k_packbits, N, 256, "packbits", grid, N, density_thresh, bitfield);
}
extern "C" int ngp_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                    float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                    uint32_t M, const float* nears, const float* fars, float* xyzs, float* dirs,
                                    float* deltas, int32_t* rays, int32_t* counter, const float* noises,
                                    ngp_stream_t stream) {
    if (C < 1 || C > 24) return fail(NGP_EINVAL, "march_rays_train: cascade count out of range");
    NGP_LAUNCH_1D(k_march_rays_train, N, MARCH_TPB, "march_rays_train", rays_o, rays_d, grid, bound, dt_gamma, max_steps, N,
                  C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises);
}
extern "C" int ngp_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                                const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                                float* weights_sum, float* depth, float* image,
                                                ngp_stream_t stream) {
    if ((uint64_t)N * 32 > 0xffffffffull) return fail(NGP_EINVAL, "composite_rays_train_forward: too many rays");
    NGP_LAUNCH_COMPOSITE(k_composite_train_fwd, "composite_rays_train_forward", sigmas, rgbs, deltas, rays, M, N,
                         T_thresh, weights_sum, depth, image, MseArgs{});
}
// Compositor + loss head in one launch (extension used by the step driver): additionally forms pred = image + (1 - ws) * bg,
// the squared error per ray (sqerr [N], row order of `rays`) and d(loss * *scale)/d(image, ws) for loss = sum(sqerr) * inv_norm / 2,
// i.e. inv_norm = 2 / (3 R) for the mean over 3 R values.  Replaces ~12 elementwise / reduction launches of the torch expression.
extern "C" int ngp_composite_rays_train_forward_mse(const float* sigmas, const float* rgbs, const float* deltas,
                                                    const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                                    const float* target, float bg, float inv_norm, const float* scale,
                                                    float* weights_sum, float* depth, float* image, float* g_image, float* g_ws,
                                                    float* sqerr, ngp_stream_t stream) {
    if ((uint64_t)N * 32 > 0xffffffffull) return fail(NGP_EINVAL, "composite_rays_train_forward_mse: too many rays");
    if (!target || !scale || !g_image || !g_ws || !sqerr) return fail(NGP_EINVAL, "composite_rays_train_forward_mse: null pointer");
    MseArgs mse{target, bg, inv_norm, scale, g_image, g_ws, sqerr};
    NGP_LAUNCH_COMPOSITE(k_composite_train_fwd, "composite_rays_train_forward_mse", sigmas, rgbs, deltas, rays, M, N,
                         T_thresh, weights_sum, depth, image, mse);
}
// ring [1] i32 (next row), counter [2] i32 (this march's totals), nsteps [1] i32, step_counter [16,2] i32
extern "C" int ngp_step_counter_push(int32_t* ring, const int32_t* counter, int32_t* nsteps, int32_t* step_counter,
                                     ngp_stream_t stream) {
    if (!ring || !counter || !nsteps || !step_counter) return fail(NGP_EINVAL, "step_counter_push: null pointer");
    k_step_counter_push<<<1, 1, 0, as_stream(stream)>>>(ring, counter, nsteps, step_counter);
    return check_launch("step_counter_push");
}
extern "C" int ngp_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                                 const float* sigmas, const float* rgbs, const float* deltas,
                                                 const int32_t* rays, const float* weights_sum, const float* image,
                                                 uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                                 float* grad_rgbs, ngp_stream_t stream) {
    if ((uint64_t)N * 32 > 0xffffffffull) return fail(NGP_EINVAL, "composite_rays_train_backward: too many rays");
    NGP_LAUNCH_COMPOSITE(k_composite_train_bwd, "composite_rays_train_backward", grad_weights_sum, grad_image, sigmas,
                         rgbs, deltas, rays, weights_sum, image, M, N, T_thresh, grad_sigmas, grad_rgbs);
}
extern "C" int ngp_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                              const float* rays_o, const float* rays_d, float bound, float dt_gamma,
                              uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, const float* nears,
                              const float* fars, float* xyzs, float* dirs, float* deltas, const float* noises,
                              ngp_stream_t stream) {
    NGP_LAUNCH_1D(k_march_rays, n_alive, 128, "march_rays", n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound,
                  dt_gamma, max_steps, C, H, grid, nears, fars, xyzs, dirs, deltas, noises);
}
extern "C" int ngp_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive,
                                  float* rays_t, const float* sigmas, const float* rgbs, const float* deltas,
                                  float* weights_sum, float* depth, float* image, ngp_stream_t stream) {
    NGP_LAUNCH_1D(k_composite_rays, n_alive, 128, "composite_rays", n_alive, n_step, T_thresh, rays_alive, rays_t,
                  sigmas, rgbs, deltas, weights_sum, depth, image);
}

// ---- device-driven inference loop (extension; see the kernels' comment) ------------------------------------------------------------
extern "C" int ngp_infer_init(uint32_t N, int32_t* rays_alive, uint32_t* ctrl, ngp_stream_t stream) {
    if (N == 0) return fail(NGP_EINVAL, "infer_init: no rays");
    NGP_LAUNCH_1D(k_infer_init, N, 256, "infer_init", N, rays_alive, ctrl);
}
extern "C" int ngp_march_rays_dev(const uint32_t* ctrl, uint32_t N, const int32_t* rays_alive, const float* rays_t,
                                  const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                                  uint32_t C, uint32_t H, const uint8_t* grid, const float* fars, float* xyzs, float* dirs,
                                  float* deltas, const float* noises, ngp_stream_t stream) {
    NGP_LAUNCH_1D(k_march_rays_dev, N, 128, "march_rays_dev", ctrl, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps,
                  C, H, grid, fars, xyzs, dirs, deltas, noises);
}
extern "C" int ngp_composite_rays_dev(const uint32_t* ctrl, uint32_t N, float T_thresh, int32_t* rays_alive, float* rays_t,
                                      const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                                      float* depth, float* image, ngp_stream_t stream) {
    NGP_LAUNCH_1D(k_composite_rays_dev, N, 128, "composite_rays_dev", ctrl, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas,
                  weights_sum, depth, image);
}
extern "C" int ngp_compact_rays_dev(uint32_t* ctrl, uint32_t N, uint32_t max_steps, const int32_t* rays_in, int32_t* rays_out,
                                    ngp_stream_t stream) {
    if (N == 0) return NGP_OK;
    k_compact_rays_dev<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(ctrl, rays_in, rays_out);
    int rc = check_launch("compact_rays_dev");
    if (rc) return rc;
    k_infer_advance<<<1, 1, 0, as_stream(stream)>>>(ctrl, N, max_steps);
    return check_launch("compact_rays_dev(advance)");
}
