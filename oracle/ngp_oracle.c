/*
 * ngp_oracle.c — CPU restatement of the torch-ngp hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may call it.  The product (torch-ngp_b200/) never does.
 *
 * Parity status: PINNED by tests/golden/ (vectors produced on a B200 by the reference's own CUDA
 * extensions, oracle/_ref, via tests/golden/make_golden.py).  The reference repo itself holds no
 * golden vectors or known-answer tests for this path.
 *
 * Each function cites the reference lines it restates.  Where nvcc contracts a*b+c into one FFMA in
 * the reference's build (checked in the SASS of oracle/_ref/raymarching), this file calls fmaf()
 * explicitly and is compiled with -ffp-contract=off, so the float recurrences agree bit-for-bit.
 * Two device functions cannot be reproduced on a CPU: exp2f (MUFU.EX2 based; used for the 16 level
 * scales of the hash grid) and __expf (compositor).  For the former the caller may pass the
 * device-computed scale table; the latter is compared within a tolerance.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -fopenmp -o liboracle.so ngp_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

typedef _Float16 h16;

/* ------------------------------------------------------------------------------------------ */
/* hash grid  (gridencoder/src/gridencoder.cu)                                                 */
/* ------------------------------------------------------------------------------------------ */

/* :50-63 fast_hash */
static uint32_t fast_hash(const uint32_t* p, uint32_t D) {
    static const uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
    for (uint32_t i = 0; i < D; ++i) r ^= p[i] * primes[i];
    return r;
}

/* :66-84 get_grid_index (without the *C + ch) */
static uint32_t grid_index(uint32_t gridtype, int align_corners, uint32_t hashmap_size, uint32_t resolution,
                           const uint32_t* p, uint32_t D) {
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; ++d) {
        index += p[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) index = fast_hash(p, D);
    return index % hashmap_size;
}

/* :138 scale = exp2f(level*S)*H - 1 (contracted to one FMA by nvcc) */
float oracle_grid_level_scale(uint32_t level, float S, uint32_t H) {
    return fmaf(exp2f((float)level * S), (float)H, -1.0f);
}

static float smoothstep(float v) { return v * v * (3.0f - 2.0f * v); }
static float smoothstep_d(float v) { return 6 * v * (1.0f - v); }

/* :87-245 kernel_grid.  table/out dtype: 0 = f32, 1 = f16.  out is [B, L*C] (the layout grid.py:57
 * hands to the caller).  scales (nullable): per-level scale override.  indices_out (nullable):
 * [B, L, 2^D] entry indices (within level) for exact index parity checks.  dy_dx (nullable) [B,L,D,C]. */
void oracle_grid_forward(const float* inputs, const void* table, const int32_t* offsets, void* out, uint32_t B,
                         uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                         int align_corners, uint32_t interp, int dtype, const float* scales, uint32_t* indices_out,
                         void* dy_dx) {
    const uint32_t NC = 1u << D;
#pragma omp parallel for schedule(static)
    for (int64_t bb = 0; bb < (int64_t)B; ++bb) {
        const uint32_t b = (uint32_t)bb;
        const float* x = inputs + (size_t)b * D;
        int oob = 0;
        for (uint32_t d = 0; d < D; ++d) if (x[d] < 0 || x[d] > 1) oob = 1;   /* :110-117 */
        for (uint32_t level = 0; level < L; ++level) {
            float* of = (float*)out + ((size_t)b * L + level) * C;
            h16* oh = (h16*)out + ((size_t)b * L + level) * C;
            float* jf = dy_dx ? (float*)dy_dx + ((size_t)b * L + level) * D * C : 0;
            h16* jh = dy_dx ? (h16*)dy_dx + ((size_t)b * L + level) * D * C : 0;
            if (oob) {                                                         /* :118-135 */
                for (uint32_t c = 0; c < C; ++c) { if (dtype) oh[c] = 0; else of[c] = 0; }
                if (dy_dx) for (uint32_t i = 0; i < D * C; ++i) { if (dtype) jh[i] = 0; else jf[i] = 0; }
                if (indices_out) for (uint32_t i = 0; i < NC; ++i) indices_out[((size_t)b * L + level) * NC + i] = 0xffffffffu;
                continue;
            }
            const uint32_t off = (uint32_t)offsets[level];
            const uint32_t hashmap_size = (uint32_t)offsets[level + 1] - off;
            const float scale = scales ? scales[level] : oracle_grid_level_scale(level, S, H);
            const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
            float pos[8], pos_deriv[8];
            uint32_t pg[8];
            for (uint32_t d = 0; d < D; ++d) {                                  /* :146-159 */
                pos[d] = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
                pg[d] = (uint32_t)floorf(pos[d]);
                pos[d] -= (float)pg[d];
                if (interp == 1) { pos_deriv[d] = smoothstep_d(pos[d]); pos[d] = smoothstep(pos[d]); }
                else pos_deriv[d] = 1.0f;
            }
            float rf[8] = {0};
            h16 rh[8] = {0};
            for (uint32_t idx = 0; idx < NC; ++idx) {                           /* :166-191 */
                float w = 1;
                uint32_t pl[8];
                for (uint32_t d = 0; d < D; ++d) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                    else { w *= pos[d]; pl[d] = pg[d] + 1; }
                }
                const uint32_t index = grid_index(gridtype, align_corners, hashmap_size, resolution, pl, D);
                if (indices_out) indices_out[((size_t)b * L + level) * NC + idx] = index;
                for (uint32_t c = 0; c < C; ++c) {
                    if (dtype) {  /* c10::Half: product rounded to half, then half += half via float (:187) */
                        const h16 g = ((const h16*)table)[((size_t)off + index) * C + c];
                        const h16 p = (h16)(w * (float)g);
                        rh[c] = (h16)((float)rh[c] + (float)p);
                    } else {
                        const float g = ((const float*)table)[((size_t)off + index) * C + c];
                        rf[c] = fmaf(w, g, rf[c]);
                    }
                }
            }
            for (uint32_t c = 0; c < C; ++c) { if (dtype) oh[c] = rh[c]; else of[c] = rf[c]; }
            if (dy_dx) {                                                        /* :201-244 */
                for (uint32_t gd = 0; gd < D; ++gd) {
                    float gf[8] = {0};
                    h16 gh[8] = {0};
                    for (uint32_t idx = 0; idx < (1u << (D - 1)); ++idx) {
                        float w = scale;
                        uint32_t pl[8];
                        for (uint32_t nd = 0; nd < D - 1; ++nd) {
                            const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                            if ((idx & (1u << nd)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                            else { w *= pos[d]; pl[d] = pg[d] + 1; }
                        }
                        pl[gd] = pg[gd];
                        const uint32_t il = grid_index(gridtype, align_corners, hashmap_size, resolution, pl, D);
                        pl[gd] = pg[gd] + 1;
                        const uint32_t ir = grid_index(gridtype, align_corners, hashmap_size, resolution, pl, D);
                        for (uint32_t c = 0; c < C; ++c) {
                            if (dtype) {
                                const h16 vl = ((const h16*)table)[((size_t)off + il) * C + c];
                                const h16 vr = ((const h16*)table)[((size_t)off + ir) * C + c];
                                const h16 diff = (h16)((float)vr - (float)vl);
                                const h16 p = (h16)(w * (float)diff * pos_deriv[gd]);
                                gh[c] = (h16)((float)gh[c] + (float)p);
                            } else {
                                const float vl = ((const float*)table)[((size_t)off + il) * C + c];
                                const float vr = ((const float*)table)[((size_t)off + ir) * C + c];
                                gf[c] = fmaf(w * (vr - vl), pos_deriv[gd], gf[c]);
                            }
                        }
                    }
                    for (uint32_t c = 0; c < C; ++c) { if (dtype) jh[gd * C + c] = gh[c]; else jf[gd * C + c] = gf[c]; }
                }
            }
        }
    }
}

/* :248-340 kernel_grid_backward.  grad is [B, L*C]; grad_table [sO, C] accumulated in DOUBLE so the
 * result is the order-independent "true" scatter-add (the device result is an fp16/fp32 atomic sum
 * in arbitrary order; tests compare within the corresponding tolerance).  The per-corner addend is
 * rounded exactly like the device does: fp16 tables -> half(w*g), fp32 tables -> float(w*g). */
void oracle_grid_backward(const void* grad, const float* inputs, const int32_t* offsets, double* grad_table,
                          uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                          int align_corners, uint32_t interp, int dtype, const float* scales) {
    const uint32_t NC = 1u << D;
    for (uint32_t b = 0; b < B; ++b) {
        const float* x = inputs + (size_t)b * D;
        int oob = 0;
        for (uint32_t d = 0; d < D; ++d) if (x[d] < 0 || x[d] > 1) oob = 1;
        if (oob) continue;
        for (uint32_t level = 0; level < L; ++level) {
            const uint32_t off = (uint32_t)offsets[level];
            const uint32_t hashmap_size = (uint32_t)offsets[level + 1] - off;
            const float scale = scales ? scales[level] : oracle_grid_level_scale(level, S, H);
            const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
            float pos[8];
            uint32_t pg[8];
            for (uint32_t d = 0; d < D; ++d) {
                pos[d] = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
                pg[d] = (uint32_t)floorf(pos[d]);
                pos[d] -= (float)pg[d];
                if (interp == 1) pos[d] = smoothstep(pos[d]);
            }
            for (uint32_t idx = 0; idx < NC; ++idx) {
                float w = 1;
                uint32_t pl[8];
                for (uint32_t d = 0; d < D; ++d) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                    else { w *= pos[d]; pl[d] = pg[d] + 1; }
                }
                const uint32_t index = grid_index(gridtype, align_corners, hashmap_size, resolution, pl, D);
                for (uint32_t c = 0; c < C; ++c) {
                    double add;
                    if (dtype) add = (double)(float)(h16)(w * (float)((const h16*)grad)[((size_t)b * L + level) * C + c]);
                    else add = (double)(w * ((const float*)grad)[((size_t)b * L + level) * C + c]);
                    grad_table[((size_t)off + index) * C + c] += add;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* ray marching  (raymarching/src/raymarching.cu)                                              */
/* ------------------------------------------------------------------------------------------ */

static float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static float signf1(float x) { return copysignf(1.0f, x); }

/* :56-72 */
static uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
uint32_t oracle_morton3D(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
/* :74-82 */
uint32_t oracle_morton3D_invert(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* :42-54 */
static int mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}
static int mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = (float)((double)(dt * H) * 0.5);
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}

/* :91-145 */
void oracle_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                               float min_near, float* nears, float* fars) {
    for (uint32_t n = 0; n < N; ++n) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float rdx = 1 / rays_d[n * 3], rdy = 1 / rays_d[n * 3 + 1], rdz = 1 / rays_d[n * 3 + 2];
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, c;
        if (near > far) { c = near; near = far; far = c; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { c = near_y; near_y = far_y; far_y = c; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { c = near_z; near_z = far_z; far_z = c; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* :267-289 */
void oracle_packbits(const float* grid, uint32_t N, float thresh, uint8_t* bitfield) {
    for (uint32_t n = 0; n < N; ++n) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; ++i) bits |= (grid[(size_t)n * 8 + i] > thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, bound, dt_gamma, dt_min, dt_max, rH, H3, fH, fC, Hm1;
    uint32_t H;
} ray_t;

static void ray_setup(ray_t* r, const float* o, const float* d, float bound, float dt_gamma, uint32_t max_steps,
                      uint32_t C, uint32_t H) {
    r->ox = o[0]; r->oy = o[1]; r->oz = o[2];
    r->dx = d[0]; r->dy = d[1]; r->dz = d[2];
    r->rdx = 1 / r->dx; r->rdy = 1 / r->dy; r->rdz = 1 / r->dz;                 /* :341 */
    r->bound = bound; r->dt_gamma = dt_gamma; r->H = H;
    r->fH = (float)H; r->fC = (float)C; r->Hm1 = (float)(H - 1);
    r->rH = 1 / (float)H;                                                       /* :342 */
    r->H3 = (float)(H * H * H);                                                 /* :343 */
    const float two_sqrt3 = 2 * 1.7320508075688772f;
    r->dt_min = two_sqrt3 / (float)max_steps;                                   /* :349 */
    r->dt_max = two_sqrt3 * (float)(1 << (C - 1)) / (float)H;                   /* :350 */
}

/* one loop body of :361-399 (identical at :426-478 and :748-803) */
static int march_probe(const ray_t* r, const uint8_t* grid, float* t, float* x, float* y, float* z, float* dt) {
    *x = clampf(fmaf(*t, r->dx, r->ox), -r->bound, r->bound);
    *y = clampf(fmaf(*t, r->dy, r->oy), -r->bound, r->bound);
    *z = clampf(fmaf(*t, r->dz, r->oz), -r->bound, r->bound);
    *dt = clampf(*t * r->dt_gamma, r->dt_min, r->dt_max);
    const int a = mip_from_pos(*x, *y, *z, r->fC), bq = mip_from_dt(*dt, r->fH, r->fC);
    const int level = a > bq ? a : bq;
    const float mip_bound = fminf(scalbnf(1.0f, level), r->bound);
    const float mip_rbound = 1 / mip_bound;
    const int nx = (int)clampf((float)(0.5 * (double)fmaf(*x, mip_rbound, 1.0f) * (double)r->H), 0.0f, r->Hm1);
    const int ny = (int)clampf((float)(0.5 * (double)fmaf(*y, mip_rbound, 1.0f) * (double)r->H), 0.0f, r->Hm1);
    const int nz = (int)clampf((float)(0.5 * (double)fmaf(*z, mip_rbound, 1.0f) * (double)r->H), 0.0f, r->Hm1);
    const uint32_t index = (uint32_t)fmaf((float)level, r->H3, (float)oracle_morton3D(nx, ny, nz));
    if (grid[index / 8] & (1 << (index % 8))) return 1;
    const float tx = fmaf(fmaf(fmaf(0.5f, signf1(r->dx), (float)nx + 0.5f) * r->rH, 2.0f, -1.0f), mip_bound, -*x) * r->rdx;
    const float ty = fmaf(fmaf(fmaf(0.5f, signf1(r->dy), (float)ny + 0.5f) * r->rH, 2.0f, -1.0f), mip_bound, -*y) * r->rdy;
    const float tz = fmaf(fmaf(fmaf(0.5f, signf1(r->dz), (float)nz + 0.5f) * r->rH, 2.0f, -1.0f), mip_bound, -*z) * r->rdz;
    const float tt = *t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    do { *t += clampf(*t * r->dt_gamma, r->dt_min, r->dt_max); } while (*t < tt);
    return 0;
}

/* :311-480.  Rays are processed in ray order, so rays[k] = (k, prefix offset, count): one valid
 * instance of the reference's atomics-ordered output.  xyzs/dirs/deltas must be zero-filled. */
void oracle_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                             float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                             const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                             int32_t* rays, int32_t* counter, const float* noises) {
    for (uint32_t n = 0; n < N; ++n) {
        ray_t r;
        ray_setup(&r, rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, bound, dt_gamma, max_steps, C, H);
        const float near = nears[n], far = fars[n];
        const float t0 = fmaf(clampf(near * dt_gamma, r.dt_min, r.dt_max), noises[n], near);   /* :355 */
        float t = t0, x, y, z, dt;
        uint32_t num_steps = 0;
        while (t < far && num_steps < max_steps)
            if (march_probe(&r, grid, &t, &x, &y, &z, &dt)) { num_steps++; t += dt; }
        const uint32_t point_index = (uint32_t)counter[0]; counter[0] += (int32_t)num_steps;    /* :405 */
        const uint32_t ray_index = (uint32_t)counter[1]; counter[1] += 1;                       /* :406 */
        rays[ray_index * 3] = (int32_t)n;
        rays[ray_index * 3 + 1] = (int32_t)point_index;
        rays[ray_index * 3 + 2] = (int32_t)num_steps;
        if (num_steps == 0) continue;
        if (point_index + num_steps > M) continue;                                              /* :416 */
        float* px = xyzs + (size_t)point_index * 3;
        float* pd = dirs + (size_t)point_index * 3;
        float* pl = deltas + (size_t)point_index * 2;
        t = t0;
        float last_t = t;
        uint32_t step = 0;
        while (t < far && step < num_steps) {
            if (march_probe(&r, grid, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

/* :500-577 (expf instead of the device's __expf) */
void oracle_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                         const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                         float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) {
            weights_sum[index] = 0; depth[index] = 0;
            image[index * 3] = image[index * 3 + 1] = image[index * 3 + 2] = 0;
            continue;
        }
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        for (uint32_t s = 0; s < num_steps; ++s) {
            const uint32_t k = offset + s;
            const float alpha = 1.0f - expf(-sigmas[k] * deltas[k * 2]);
            const float weight = alpha * T;
            r = fmaf(weight, rgbs[k * 3], r); g = fmaf(weight, rgbs[k * 3 + 1], g); b = fmaf(weight, rgbs[k * 3 + 2], b);
            t += deltas[k * 2 + 1];
            d = fmaf(weight, t, d);
            ws += weight;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
        }
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* :601-682.  grad_sigmas / grad_rgbs must be zero-filled (raymarching.py:283-284). */
void oracle_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                          const float* sigmas, const float* rgbs, const float* deltas,
                                          const int32_t* rays, const float* weights_sum, const float* image,
                                          uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                          float* grad_rgbs) {
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) continue;
        const float gws = grad_weights_sum[index];
        const float* gi = grad_image + (size_t)index * 3;
        const float rf = image[index * 3], gf = image[index * 3 + 1], bf = image[index * 3 + 2], wsf = weights_sum[index];
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
        for (uint32_t s = 0; s < num_steps; ++s) {
            const uint32_t k = offset + s;
            const float alpha = 1.0f - expf(-sigmas[k] * deltas[k * 2]);
            const float weight = alpha * T;
            r = fmaf(weight, rgbs[k * 3], r); g = fmaf(weight, rgbs[k * 3 + 1], g); b = fmaf(weight, rgbs[k * 3 + 2], b);
            ws += weight;
            T *= 1.0f - alpha;
            grad_rgbs[k * 3] = gi[0] * weight; grad_rgbs[k * 3 + 1] = gi[1] * weight; grad_rgbs[k * 3 + 2] = gi[2] * weight;
            grad_sigmas[k] = deltas[k * 2] * (gi[0] * (T * rgbs[k * 3] - (rf - r)) + gi[1] * (T * rgbs[k * 3 + 1] - (gf - g)) +
                                              gi[2] * (T * rgbs[k * 3 + 2] - (bf - b)) + gws * (1 - wsf));
            if (T < T_thresh) break;
        }
    }
}

/* :700-805.  xyzs/dirs/deltas must be zero-filled (raymarching.py:333-335). */
void oracle_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                       const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                       uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars,
                       float* xyzs, float* dirs, float* deltas, const float* noises) {
    (void)nears;
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int index = rays_alive[n];
        ray_t r;
        ray_setup(&r, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, bound, dt_gamma, max_steps, C, H);
        float* px = xyzs + (size_t)n * n_step * 3;
        float* pd = dirs + (size_t)n * n_step * 3;
        float* pl = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index];
        const float far = fars[index];
        t = fmaf(clampf(t * dt_gamma, r.dt_min, r.dt_max), noises[n], t);      /* :744 */
        float last_t = t, x, y, z, dt;
        uint32_t step = 0;
        while (t < far && step < n_step) {
            if (march_probe(&r, grid, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

/* :818-905 */
void oracle_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                           const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                           float* depth, float* image) {
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int index = rays_alive[n];
        const float* sg = sigmas + (size_t)n * n_step;
        const float* cl = rgbs + (size_t)n * n_step * 3;
        const float* dl = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index], weight_sum = weights_sum[index], d = depth[index];
        float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        uint32_t step = 0;
        while (step < n_step) {
            if (dl[step * 2] == 0) break;
            const float alpha = 1.0f - expf(-sg[step] * dl[step * 2]);
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            t += dl[step * 2 + 1];
            d = fmaf(weight, t, d);
            r = fmaf(weight, cl[step * 3], r); g = fmaf(weight, cl[step * 3 + 1], g); b = fmaf(weight, cl[step * 3 + 2], b);
            if (T < T_thresh) break;
            step++;
        }
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        weights_sum[index] = weight_sum; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* occupancy-grid maintenance (nerf/renderer.py — torch op sequences restated on scalars)       */
/* ------------------------------------------------------------------------------------------ */
/* mark_untrained_grid, renderer.py:380-442.  One cell at a time: world = (2c/(H-1) - 1) * (bound_c - hgs) (:413-419, every
 * torch op rounds separately; tensor/scalar division on a CUDA device multiplies by the fp32 reciprocal); cam = (world - t) @ R
 * (:428-429, accumulated with fused multiply-adds in index order — cuBLAS does not document its order, the comparison with
 * the reference's own output is therefore a count of boundary cells, see tests); masks :432-435; count==0 -> -1 (:440). */
void oracle_mark_untrained(const float* poses, uint32_t B, double fx, double fy, double cx, double cy, float bound,
                           uint32_t C, uint32_t H, float* grid, uint32_t* count_out) {
    const uint32_t H3 = H * H * H;
    const float inv = 1.0f / (float)(H - 1);
    const float rx = (float)(cx / fx), ry = (float)(cy / fy);
    for (uint32_t cas = 0; cas < C; ++cas) {
        const double bc = fmin((double)(1u << cas), (double)bound);
        const double hg = bc / (double)H;
        const float s = (float)(bc - hg), hgs2 = (float)(hg * 2.0);
#pragma omp parallel for
        for (uint32_t i = 0; i < H3; ++i) {
            float w[3];
            for (int d = 0; d < 3; ++d) {
                const uint32_t c = oracle_morton3D_invert(i >> d);
                w[d] = ((2.0f * (float)c) * inv - 1.0f) * s;
            }
            uint32_t count = 0;
            for (uint32_t b = 0; b < B; ++b) {
                const float* P = poses + (size_t)b * 16;
                const float dx = w[0] - P[3], dy = w[1] - P[7], dz = w[2] - P[11];
                const float cxx = fmaf(dz, P[8], fmaf(dy, P[4], dx * P[0]));
                const float cyy = fmaf(dz, P[9], fmaf(dy, P[5], dx * P[1]));
                const float czz = fmaf(dz, P[10], fmaf(dy, P[6], dx * P[2]));
                if (czz > 0.0f && fabsf(cxx) < rx * czz + hgs2 && fabsf(cyy) < ry * czz + hgs2) ++count;
            }
            if (count_out) count_out[(size_t)cas * H3 + i] = count;
            if (count == 0) grid[(size_t)cas * H3 + i] = -1.0f;
        }
    }
}
