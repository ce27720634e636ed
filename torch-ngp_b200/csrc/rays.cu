// rays.cu — ray generation + training-pixel gather on the device (input side of the hot path).
//
// The reference builds a step's rays with ~20 torch ops (nerf/utils.py:54-137 get_rays: two H*W meshgrids, gathers, stack, norm,
// batched matmul, expand) and fetches the target pixels with a torch.gather over a stacked index (nerf/provider.py:308-312),
// then converts / alpha-blends them in the trainer (nerf/utils.py:494-508).  Here one kernel maps a pixel index to
// (rays_o, rays_d) and one gathers + (optionally) linearises + blends the target colour: 36 B written and 12-16 B read per ray,
// no H*W-sized temporaries.
//
// Arithmetic restates the torch ops in order (tensor/scalar division as multiplication by the fp32 reciprocal, like torch's
// CUDA kernels); the norm and the 3x3 matmul are accumulated in index order with FMAs — torch.norm / cuBLAS do not document
// their order, so parity with the reference is a floating-point tolerance (tests: 1e-6), not bit-exactness.
#include "common.cuh"

namespace ngp {

static constexpr uint32_t RAYS_TPB = 256;

// poses [B,4,4] c2w; inds [*, N] pixel indices with batch stride inds_stride (0 = shared by all cameras, the reference's
// inds.expand([B, N])) or NULL (ray n = pixel n)
__global__ void __launch_bounds__(RAYS_TPB) k_get_rays(const float* __restrict__ poses, uint32_t B, uint32_t N, uint32_t W,
                                                       float cx, float cy, float inv_fx, float inv_fy,
                                                       const int64_t* __restrict__ inds, uint32_t inds_stride,
                                                       float* __restrict__ rays_o, float* __restrict__ rays_d) {
    const uint32_t t = blockIdx.x * RAYS_TPB + threadIdx.x;
    if (t >= B * N) return;
    const uint32_t b = t / N, n = t % N;
    const int64_t pix = inds ? __ldg(inds + (size_t)b * inds_stride + n) : (int64_t)n;
    const float i = __fadd_rn((float)(pix % W), 0.5f), j = __fadd_rn((float)(pix / W), 0.5f);    // utils.py:72-74
    const float xs = __fmul_rn(__fsub_rn(i, cx), inv_fx), ys = __fmul_rn(__fsub_rn(j, cy), inv_fy);   // :126-128
    const float nrm = sqrtf(fmaf(ys, ys, fmaf(xs, xs, 0.0f)) + 1.0f);                                // :130
    const float dx = __fdiv_rn(xs, nrm), dy = __fdiv_rn(ys, nrm), dz = __fdiv_rn(1.0f, nrm);
    const float* P = poses + (size_t)b * 16;
    float* o = rays_o + (size_t)t * 3;
    float* d = rays_d + (size_t)t * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        d[r] = fmaf(dz, __ldg(P + 4 * r + 2), fmaf(dy, __ldg(P + 4 * r + 1), __fmul_rn(dx, __ldg(P + 4 * r))));   // :131  d @ R^T
        o[r] = __ldg(P + 4 * r + 3);                                                                             // :133-134
    }
}

__device__ __forceinline__ float srgb_to_linear(float x) {      // utils.py:48-50
    return x < 0.04045f ? __fmul_rn(x, 1.0f / 12.92f) : powf(__fmul_rn(__fadd_rn(x, 0.055f), 1.0f / 1.055f), 2.4f);
}

// images [n_img, H*W, C] (C = 3 or 4; float32 or uint8/255), image_index [B] selects the image of each camera.
// pixels_out [B,N,C] (nullable) = the raw gather of provider.py:311; gt_out [B,N,3] (nullable) = trainer's target colour:
// optional sRGB->linear on rgb, then rgb * a + bg * (1 - a) for C == 4 (utils.py:494-508), bg = bg_pixels [B,N,3] or bg_scalar.
template <typename T>
__device__ __forceinline__ float px(const T* p);
template <> __device__ __forceinline__ float px<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float px<uint8_t>(const uint8_t* p) { return __fdiv_rn((float)__ldg(p), 255.0f); }

template <typename T>
__global__ void __launch_bounds__(RAYS_TPB) k_gather_pixels(const T* __restrict__ images, const int64_t* __restrict__ image_index,
                                                            uint32_t HW, uint32_t C, uint32_t B, uint32_t N,
                                                            const int64_t* __restrict__ inds, uint32_t inds_stride, int linear,
                                                            const float* __restrict__ bg_pixels, float bg_scalar,
                                                            float* __restrict__ pixels_out, float* __restrict__ gt_out) {
    const uint32_t t = blockIdx.x * RAYS_TPB + threadIdx.x;
    if (t >= B * N) return;
    const uint32_t b = t / N, n = t % N;
    const int64_t pix = inds ? __ldg(inds + (size_t)b * inds_stride + n) : (int64_t)n;
    const int64_t img = image_index ? __ldg(image_index + b) : (int64_t)b;
    const T* src = images + ((size_t)img * HW + (size_t)pix) * C;
    float v[4] = {0.f, 0.f, 0.f, 1.f};
    for (uint32_t c = 0; c < C; ++c) v[c] = px<T>(src + c);
    if (pixels_out)
        for (uint32_t c = 0; c < C; ++c) pixels_out[(size_t)t * C + c] = v[c];
    if (gt_out) {
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) {
            float x = linear ? srgb_to_linear(v[c]) : v[c];
            if (C == 4) {
                const float bg = bg_pixels ? __ldg(bg_pixels + (size_t)t * 3 + c) : bg_scalar;
                x = __fadd_rn(__fmul_rn(x, v[3]), __fmul_rn(bg, __fsub_rn(1.0f, v[3])));
            }
            gt_out[(size_t)t * 3 + c] = x;
        }
    }
}

}  // namespace ngp

using namespace ngp;

extern "C" int ngp_get_rays(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                            uint32_t N, const int64_t* inds, uint32_t inds_stride, float* rays_o, float* rays_d,
                            ngp_stream_t stream) {
    if (H == 0 || W == 0) return fail(NGP_EINVAL, "get_rays: empty image");
    if (!inds && N != H * W) return fail(NGP_EINVAL, "get_rays: without indices N must be H*W");
    if ((uint64_t)B * N > 0x7fffffffull) return fail(NGP_EINVAL, "get_rays: too many rays");
    if ((uint64_t)B * N == 0) return NGP_OK;
    k_get_rays<<<div_up(B * N, RAYS_TPB), RAYS_TPB, 0, as_stream(stream)>>>(poses, B, N, W, cx, cy, 1.0f / fx, 1.0f / fy, inds, inds_stride,
                                                                           rays_o, rays_d);
    return check_launch("get_rays");
}

extern "C" int ngp_gather_pixels(const void* images, int dtype, const int64_t* image_index, uint32_t H, uint32_t W, uint32_t C,
                                 uint32_t B, uint32_t N, const int64_t* inds, uint32_t inds_stride, int linear,
                                 const float* bg_pixels, float bg_scalar, float* pixels_out, float* gt_out, ngp_stream_t stream) {
    if (C != 3 && C != 4) return fail(NGP_EINVAL, "gather_pixels: images must have 3 or 4 channels");
    if (!inds && N != H * W) return fail(NGP_EINVAL, "gather_pixels: without indices N must be H*W");
    if ((uint64_t)B * N > 0x7fffffffull) return fail(NGP_EINVAL, "gather_pixels: too many rays");
    if ((uint64_t)B * N == 0) return NGP_OK;
    const uint32_t blocks = div_up(B * N, RAYS_TPB);
    if (dtype == 0)
        k_gather_pixels<float><<<blocks, RAYS_TPB, 0, as_stream(stream)>>>(static_cast<const float*>(images), image_index, H * W, C, B, N, inds,
                                                                           inds_stride, linear, bg_pixels, bg_scalar, pixels_out, gt_out);
    else if (dtype == 2)
        k_gather_pixels<uint8_t><<<blocks, RAYS_TPB, 0, as_stream(stream)>>>(static_cast<const uint8_t*>(images), image_index, H * W, C, B, N,
                                                                             inds, inds_stride, linear, bg_pixels, bg_scalar, pixels_out,
                                                                             gt_out);
    else
        return fail(NGP_EINVAL, "gather_pixels: dtype must be 0 (float32) or 2 (uint8)");
    return check_launch("gather_pixels");
}
