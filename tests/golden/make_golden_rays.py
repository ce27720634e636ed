"""Generate tests/golden/rays.npz with the reference's own get_rays / srgb_to_linear (nerf/utils.py, imported unmodified from
/root/reference and run on CPU tensors; its heavyweight import-time dependencies are stubbed, none of them is on this path).
Run in the build container:  python tests/golden/make_golden_rays.py"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "torch-ngp_b200")]
import ngp_synth   # noqa: E402


def import_reference_utils():
    for m in ("trimesh", "mcubes", "tensorboardX", "lpips", "imageio", "cv2", "matplotlib", "matplotlib.pyplot", "pandas"):
        sys.modules.setdefault(m, types.ModuleType(m))
    te = types.ModuleType("torch_ema"); te.ExponentialMovingAverage = object
    sys.modules.setdefault("torch_ema", te)
    tm = types.ModuleType("torchmetrics"); tmf = types.ModuleType("torchmetrics.functional")
    tmf.structural_similarity_index_measure = None
    sys.modules.setdefault("torchmetrics", tm); sys.modules.setdefault("torchmetrics.functional", tmf)
    sys.path.append("/root/reference")
    import importlib
    return importlib.import_module("nerf.utils")


def main():
    U = import_reference_utils()
    out = {}
    H, W = 20, 24
    poses = ngp_synth.make_cameras(3, seed=2).float()
    intr = np.array([31.7, 29.3, W / 2 + 0.25, H / 2 - 0.5])
    out["poses"], out["intrinsics"], out["H"], out["W"] = poses.numpy(), intr, H, W
    r = U.get_rays(poses, intr, H, W, -1)
    out["all_o"], out["all_d"] = r["rays_o"].numpy(), r["rays_d"].numpy()
    torch.manual_seed(7)
    r = U.get_rays(poses[:1], intr, H, W, 50)
    out["rand_inds"], out["rand_o"], out["rand_d"] = r["inds"].numpy(), r["rays_o"].numpy(), r["rays_d"].numpy()
    torch.manual_seed(8)
    r = U.get_rays(poses[:1], intr, H, W, 64, patch_size=4)
    out["patch_inds"], out["patch_d"] = r["inds"].numpy(), r["rays_d"].numpy()
    torch.manual_seed(9)
    em = torch.rand(2, 128 * 128)
    r = U.get_rays(poses[:2], intr, H, W, 40, error_map=em)
    out["err_map"], out["err_inds"], out["err_coarse"], out["err_d"] = em.numpy(), r["inds"].numpy(), r["inds_coarse"].numpy(), r["rays_d"].numpy()
    # target pixels: provider.py:311 gather + utils.py:494-508
    torch.manual_seed(10)
    images = torch.rand(3, H, W, 4)
    inds = out["rand_inds"]
    index = [2]
    px = torch.gather(images[index].view(1, -1, 4), 1, torch.stack(4 * [torch.from_numpy(inds)], -1))
    out["images"], out["px"] = images.numpy(), px.numpy()
    lin = px.clone()
    lin[..., :3] = U.srgb_to_linear(lin[..., :3])
    bg = torch.rand_like(lin[..., :3])
    out["bg"] = bg.numpy()
    out["gt_linear_bg"] = (lin[..., :3] * lin[..., 3:] + bg * (1 - lin[..., 3:])).numpy()
    out["gt_srgb_white"] = (px[..., :3] * px[..., 3:] + 1 * (1 - px[..., 3:])).numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rays.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
