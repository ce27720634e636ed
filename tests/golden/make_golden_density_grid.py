"""Generate tests/golden/density_grid.npz by running the REFERENCE's own occupancy-grid maintenance —
NeRFRenderer.mark_untrained_grid / update_extra_state, imported unmodified from /root/reference/nerf/renderer.py — on CPU
tensors.  Runs in the build container (no GPU needed):   python tests/golden/make_golden_density_grid.py

The reference code calls three integer ops of its CUDA extension (raymarching.morton3D / morton3D_invert / packbits); a
numpy stand-in with the same semantics is installed for them (those semantics are pinned bit-exactly against the real
extension by tests/golden/raymarching.npz).  Everything else — sampling, scatter, EMA, mean, threshold — is the reference's
Python, executed as is.  The fixture records the random draws the reference consumed (replayed from the same seed in the
same order), the positions it queried, the densities an analytic field returned, and the grid / bitfield / mean after
every call.  grid_size is lowered from 128 to 16 to keep the file small (it is a plain attribute of the reference class).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200")]
from oracle import density_grid_oracle as DG   # noqa: E402
from oracle import oracle as O                 # noqa: E402
import ngp_synth                               # noqa: E402

H, BOUND, SEED = 16, 2, 1234


def install_stubs():
    rm = types.ModuleType("raymarching")
    rm.morton3D = lambda c: torch.from_numpy(DG.morton3d(*[c.numpy().astype(np.uint32)[:, k] for k in range(3)]).astype(np.int32))
    rm.morton3D_invert = lambda i: torch.from_numpy(DG.morton3d_invert(i.numpy().astype(np.uint32)).astype(np.int32))

    def packbits(grid, thresh, bitfield=None):
        return torch.from_numpy(O.packbits(grid.numpy(), float(thresh)))
    rm.packbits = packbits
    sys.modules["raymarching"] = rm
    for m in ("trimesh", "pysdf", "mcubes", "tensorboardX", "lpips", "torch_ema", "torchmetrics", "imageio", "matplotlib",
              "matplotlib.pyplot", "cv2"):
        sys.modules.setdefault(m, types.ModuleType(m))
    utils_stub = types.ModuleType("nerf.utils")
    utils_stub.custom_meshgrid = lambda *a: torch.meshgrid(*a, indexing="ij")
    sys.modules["nerf.utils"] = utils_stub
    sys.path.append("/root/reference")


def blob_density(x):
    """analytic field: three Gaussian blobs (peak ~40) on a faint floor, so that some cells pass the threshold and most do not."""
    c = torch.tensor([[0.3, -0.2, 0.1], [-0.5, 0.4, -0.3], [1.2, 1.0, -0.8]])
    w = torch.tensor([40.0, 25.0, 10.0])
    r = torch.tensor([0.08, 0.05, 0.3])
    d2 = ((x[:, None, :] - c[None]) ** 2).sum(-1)
    return (w * torch.exp(-d2 / r)).sum(-1) + 1e-4


def main():
    install_stubs()
    import importlib
    R = importlib.import_module("nerf.renderer")

    log = []

    class Field(R.NeRFRenderer):
        def density(self, x):
            s = blob_density(x)
            log.append((x.clone(), s.clone()))
            return {"sigma": s}

    m = Field(bound=BOUND, cuda_ray=True, density_scale=1, density_thresh=0.01)
    m.grid_size = H
    m.density_grid = torch.zeros(m.cascade, H ** 3)
    m.density_bitfield = torch.zeros(m.cascade * H ** 3 // 8, dtype=torch.uint8)
    C = m.cascade
    out = dict(H=H, bound=BOUND, C=C, density_scale=1.0, density_thresh=0.01, decay=0.95)

    # ---- mark_untrained_grid: 6 cameras on one side of the scene, so that part of the volume is never seen
    poses = ngp_synth.make_cameras(6, radius=2.5, seed=5).float()
    intr = np.array(ngp_synth.intrinsics(64, 64), dtype=np.float64)
    m.mark_untrained_grid(poses, intr)
    out["poses"], out["intrinsic"] = poses.numpy(), intr
    out["marked_grid"] = m.density_grid.numpy().copy()

    # ---- update_extra_state: two full updates, then (iter_density forced to 16) two partial updates
    N = H ** 3 // 4
    for it in range(4):
        if it == 2:
            m.iter_density = 16
        full = m.iter_density < 16
        torch.manual_seed(SEED + it)
        log.clear()
        grid_before = m.density_grid.numpy().copy()
        m.update_extra_state()
        # replay the draws the reference just consumed (same seed, same order, same shapes)
        torch.manual_seed(SEED + it)
        if full:
            noise = torch.stack([torch.rand(H ** 3, 3) for _ in range(C)]).numpy()
            out[f"u{it}_noise"] = noise
        else:
            coords, picks, noise = [], [], []
            for cas in range(C):
                coords.append(torch.randint(0, H, (N, 3)).numpy().astype(np.int32))
                nz = int((grid_before[cas] > 0).sum())
                picks.append(torch.randint(0, nz, [N], dtype=torch.long).numpy())
                noise.append(torch.rand(2 * N, 3).numpy())
            out[f"u{it}_coords"], out[f"u{it}_picks"], out[f"u{it}_noise"] = np.stack(coords), np.stack(picks), np.stack(noise)
        out[f"u{it}_full"] = full
        out[f"u{it}_xyzs"] = np.stack([x.numpy() for x, _ in log])        # [C, n, 3] in the reference's order
        out[f"u{it}_sigmas"] = np.stack([s.numpy() for _, s in log])
        out[f"u{it}_grid"] = m.density_grid.numpy().copy()
        out[f"u{it}_bitfield"] = m.density_bitfield.numpy().copy()
        out[f"u{it}_mean"] = np.float32(m.mean_density)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "density_grid.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
