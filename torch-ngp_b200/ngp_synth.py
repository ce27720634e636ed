"""ngp_synth.py — deterministic synthetic inputs for the hot path (SURVEY §8d "Concrete synthetic inputs").

Blender-convention cameras on a sphere, 800x800 pinhole rays, a box-union occupancy grid in the
reference's on-device format (Morton order, 8 cells per byte), random-init field.  Used by bench.py,
tests/ and __graft_entry__.smoke(); contains no kernels and no oracle code.
"""
import math

import numpy as np
import torch


def _part1by2(v):
    v = v.astype(np.uint32) & 0x3ff
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def morton3d_np(x, y, z):
    return _part1by2(x) | (_part1by2(y) << 1) | (_part1by2(z) << 2)


def make_cameras(n=100, radius=4.03 * 0.8, seed=11):
    """cam2world poses [n,4,4] looking at the origin; camera +z is the viewing direction (ngp convention)."""
    rng = np.random.default_rng(seed)
    poses = np.zeros((n, 4, 4), dtype=np.float32)
    for i in range(n):
        theta = rng.uniform(0, 2 * math.pi)
        phi = math.acos(rng.uniform(-0.2, 0.95))   # mostly upper hemisphere, like blender scenes
        pos = radius * np.array([math.sin(phi) * math.cos(theta), math.cos(phi), math.sin(phi) * math.sin(theta)])
        fwd = -pos / np.linalg.norm(pos)
        up = np.array([0.0, 1.0, 0.0])
        right = np.cross(up, fwd); right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        poses[i, :3, 0] = right
        poses[i, :3, 1] = down
        poses[i, :3, 2] = fwd
        poses[i, :3, 3] = pos
        poses[i, 3, 3] = 1
    return torch.from_numpy(poses)


def intrinsics(H=800, W=800, camera_angle_x=0.6911112070083618):
    f = 0.5 * W / math.tan(0.5 * camera_angle_x)
    return (f, f, W / 2, H / 2)


def get_rays(pose, intr, H, W, inds=None, device="cpu"):
    """rays_o, rays_d [N,3] for one camera (pixel-centre rays, normalised; cf. reference nerf/utils.py:54-137)."""
    fx, fy, cx, cy = intr
    pose = pose.to(device)
    if inds is None:
        inds = torch.arange(H * W, device=device)
    inds = inds.to(device)
    j = torch.div(inds, W, rounding_mode="floor").float() + 0.5
    i = (inds % W).float() + 0.5
    dirs = torch.stack([(i - cx) / fx, (j - cy) / fy, torch.ones_like(i)], dim=-1)
    dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
    rays_d = dirs @ pose[:3, :3].T
    rays_o = pose[:3, 3].expand_as(rays_d)
    return rays_o.contiguous(), rays_d.contiguous()


def box_union_density(H=128, n_boxes=24, extent=0.7, fill_target=0.05, seed=12):
    """Density grid float32 [1, H^3] in Morton order: 1.0 inside a union of axis-aligned boxes that fills about
    `fill_target` of [-extent, extent]^3 (a 'lego-shaped' synthetic occupancy), 0 elsewhere."""
    rng = np.random.default_rng(seed)
    occ = np.zeros((H, H, H), dtype=bool)
    lo_c = int((1 - extent) / 2 * H); hi_c = int((1 + extent) / 2 * H)
    vol = (hi_c - lo_c) ** 3
    tries = 0
    while occ.sum() < fill_target * vol and tries < 10 * n_boxes:
        size = rng.integers(max(2, H // 32), max(3, H // 6), size=3)
        lo = np.array([rng.integers(lo_c, max(lo_c + 1, hi_c - s)) for s in size])
        occ[lo[0]:lo[0] + size[0], lo[1]:lo[1] + size[1], lo[2]:lo[2] + size[2]] = True
        tries += 1
    xs, ys, zs = np.nonzero(occ)
    grid = np.zeros(H ** 3, dtype=np.float32)
    grid[morton3d_np(xs, ys, zs)] = 1.0
    return torch.from_numpy(grid).view(1, -1), float(occ.mean())


def packbits_np(grid, thresh=0.01):
    g = (grid.reshape(-1, 8) > thresh)
    w = (1 << np.arange(8)).astype(np.uint8)
    return (g * w).sum(axis=1).astype(np.uint8)
