"""rays_oracle.py — CPU restatement of the reference's ray generation and target-pixel preparation.  TEST INFRASTRUCTURE ONLY.

    get_rays       nerf/utils.py:54-137   (per-pixel arithmetic :72-134; the pixel choice is an input)
    gather_pixels  nerf/provider.py:308-312 (torch.gather) + nerf/utils.py:494-508 (srgb_to_linear :48-50, alpha blend)

numpy float32, one rounding per elementwise op like the torch ops restated; the norm and the 3x3 product are evaluated in float64
and rounded once (torch.norm / cuBLAS accumulate in an undocumented order — the comparison is a 1e-6 tolerance either way).
Parity status: PINNED by tests/golden/rays.npz, produced by the reference's own get_rays / srgb_to_linear imported unmodified
from /root/reference/nerf/utils.py and run on CPU (tests/golden/make_golden_rays.py).
"""
import numpy as np

f32 = np.float32


def get_rays(poses, intrinsics, H, W, inds=None):
    poses = np.asarray(poses, f32).reshape(-1, 4, 4)
    B = poses.shape[0]
    fx, fy, cx, cy = [f32(v) for v in intrinsics]
    if inds is None:
        inds = np.arange(H * W)[None].repeat(B, 0)
    inds = np.broadcast_to(np.asarray(inds, np.int64).reshape(-1, np.asarray(inds).shape[-1]), (B, np.asarray(inds).shape[-1]))
    i = (inds % W).astype(f32) + f32(0.5)
    j = (inds // W).astype(f32) + f32(0.5)
    xs = (i - cx) * (f32(1) / fx)
    ys = (j - cy) * (f32(1) / fy)
    d = np.stack([xs, ys, np.ones_like(xs)], -1).astype(np.float64)
    d = (d / np.sqrt((d * d).sum(-1, keepdims=True)).astype(f32).astype(np.float64)).astype(f32)
    rays_d = np.einsum('bnc,bkc->bnk', d.astype(np.float64), poses[:, :3, :3].astype(np.float64)).astype(f32)
    rays_o = np.broadcast_to(poses[:, None, :3, 3], rays_d.shape).copy()
    return rays_o, rays_d


def srgb_to_linear(x):
    x = np.asarray(x, f32)
    return np.where(x < f32(0.04045), x / f32(12.92), (((x + f32(0.055)) / f32(1.055)).astype(np.float64) ** 2.4).astype(f32))


def gather_pixels(images, inds, image_index=None, gt=False, linear=False, bg=1.0):
    images = np.asarray(images)
    if images.dtype == np.uint8:
        images = images.astype(f32) / f32(255)
    n_img, H, W, C = images.shape
    inds = np.asarray(inds, np.int64)
    inds = inds.reshape(-1, inds.shape[-1])
    B = inds.shape[0] if image_index is None else len(image_index)
    inds = np.broadcast_to(inds, (B, inds.shape[-1]))
    sel = np.arange(B) if image_index is None else np.asarray(image_index)
    px = images.reshape(n_img, H * W, C)[sel[:, None], inds]
    if not gt:
        return px
    rgb = srgb_to_linear(px[..., :3]) if linear else px[..., :3]
    if C == 4:
        a = px[..., 3:]
        rgb = rgb * a + np.asarray(bg, f32) * (f32(1) - a)
    return rgb.astype(f32)
