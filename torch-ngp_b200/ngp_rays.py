"""ngp_rays.py — ray generation and training-pixel gather on the device.

`get_rays` mirrors nerf/utils.py:54-137 of the reference (same arguments, same result dict, the same torch RNG calls for the
pixel choice); the per-ray arithmetic — what the reference spreads over two H*W meshgrids, gathers, stack, norm, matmul and
expand — is one kernel (ngp_get_rays).  `gather_pixels` is provider.py:308-312's torch.gather plus, optionally, the trainer's
colour-space conversion and alpha blend (nerf/utils.py:494-508) in the same pass (ngp_gather_pixels).  No CPU path.
"""
import torch

import _ngp_b200 as _backend


def _inds_arg(inds, B, N):
    if inds is None:
        return None, 0
    if inds.dim() == 1:
        inds = inds.view(1, -1)
    if inds.shape[-1] != N:
        raise RuntimeError("ngp_rays: index list length mismatch")
    shared = inds.shape[0] == 1 or inds.stride(0) == 0
    inds = (inds[:1] if shared else inds).long().contiguous()
    if not shared and inds.shape[0] != B:
        raise RuntimeError("ngp_rays: per-camera index lists must have one row per camera")
    return inds, (0 if shared else N)


def rays_from_pixels(poses, intrinsics, H, W, inds=None):
    """rays_o, rays_d [B,N,3] for pixel indices inds ([N], [1,N] or [B,N]; None = every pixel)."""
    _backend.require_cuda(poses)
    poses = poses.float().contiguous().view(-1, 4, 4)
    B = poses.shape[0]
    N = H * W if inds is None else inds.shape[-1]
    inds_c, stride = _inds_arg(inds, B, N)
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    rays_o = torch.empty(B, N, 3, dtype=torch.float32, device=poses.device)
    rays_d = torch.empty(B, N, 3, dtype=torch.float32, device=poses.device)
    _backend.call("ngp_get_rays", poses.data_ptr(), B, fx, fy, cx, cy, H, W, N, _backend.ptr(inds_c), stride, rays_o.data_ptr(),
                  rays_d.data_ptr())
    return rays_o, rays_d


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, error_map=None, patch_size=1):
    """Drop-in for nerf/utils.py:54-137 (the pixel choice consumes torch's RNG exactly as the reference does)."""
    device = poses.device
    B = poses.shape[0]
    results = {}
    inds = None
    if N > 0:
        N = min(N, H * W)
        if patch_size > 1:
            num_patch = N // (patch_size ** 2)
            top = torch.randint(0, H - patch_size, size=[num_patch], device=device)
            left = torch.randint(0, W - patch_size, size=[num_patch], device=device)
            off = torch.arange(patch_size, device=device)
            rows = (top[:, None, None] + off[None, :, None]).expand(num_patch, patch_size, patch_size)
            cols = (left[:, None, None] + off[None, None, :]).expand(num_patch, patch_size, patch_size)
            inds = (rows * W + cols).reshape(-1)
            inds = inds.expand([B, inds.shape[0]])
        elif error_map is None:
            inds = torch.randint(0, H * W, size=[N], device=device).expand([B, N])
        else:
            coarse = torch.multinomial(error_map.to(device), N, replacement=False)
            cr, cc = coarse // 128, coarse % 128
            sx, sy = H / 128, W / 128
            r = (cr * sx + torch.rand(B, N, device=device) * sx).long().clamp(max=H - 1)
            c = (cc * sy + torch.rand(B, N, device=device) * sy).long().clamp(max=W - 1)
            inds = r * W + c
            results['inds_coarse'] = coarse
        results['inds'] = inds
    results['rays_o'], results['rays_d'] = rays_from_pixels(poses, intrinsics, H, W, inds)
    return results


@torch.no_grad()
def gather_pixels(images, inds, image_index=None, gt=False, linear=False, bg_color=1.0):
    """images [n_img,H,W,C] float32 or uint8 on the device; inds [B,N] (or [N]).  gt=False: the raw [B,N,C] gather of
    provider.py:311.  gt=True: the trainer's target [B,N,3] (optional sRGB->linear, alpha blend over bg_color: a float or a
    [B,N,3] tensor)."""
    _backend.require_cuda(images)
    if images.dtype not in (torch.float32, torch.uint8):
        raise RuntimeError("gather_pixels: images must be float32 or uint8")
    images = images.contiguous()
    n_img, H, W, C = images.shape
    if inds.dim() == 1:
        inds = inds.view(1, -1)
    N = inds.shape[-1]
    B = inds.shape[0] if image_index is None else len(image_index)
    inds_c, stride = _inds_arg(inds, B, N)
    idx = None if image_index is None else torch.as_tensor(image_index, dtype=torch.long, device=images.device).contiguous()
    if idx is None and B > n_img:
        raise RuntimeError("gather_pixels: more cameras than images")
    out_p = None if gt else torch.empty(B, N, C, dtype=torch.float32, device=images.device)
    out_g = torch.empty(B, N, 3, dtype=torch.float32, device=images.device) if gt else None
    bg_t = bg_color.float().contiguous() if torch.is_tensor(bg_color) else None
    _backend.call("ngp_gather_pixels", images.data_ptr(), 0 if images.dtype == torch.float32 else 2, _backend.ptr(idx), H, W, C, B, N,
                  _backend.ptr(inds_c), stride, int(linear), _backend.ptr(bg_t), 1.0 if bg_t is not None else float(bg_color),
                  _backend.ptr(out_p), _backend.ptr(out_g))
    return out_g if gt else out_p
