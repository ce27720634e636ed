"""density_grid_oracle.py — CPU restatement of the reference's occupancy-grid maintenance.  TEST INFRASTRUCTURE ONLY
(only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it; the product never does).

Follows nerf/renderer.py of the reference line by line on numpy float32 arrays (numpy rounds every elementwise op
separately, like the torch ops it restates):
    mark_untrained_grid  :380-442   (C, needs fmaf — oracle_mark_untrained in ngp_oracle.c)
    update_extra_state   :445-538   (sample_full :456-483, occupied :495, sample_partial :487-509, update :511-530)
The one deliberate difference from running the reference on a CPU: `tensor / python_scalar` is restated as multiplication by the
fp32 reciprocal, which is what torch's CUDA division kernel computes (the reference only ever runs this code on a GPU).

Parity status: PINNED by tests/golden/density_grid.npz — produced by the reference's own NeRFRenderer.mark_untrained_grid /
update_extra_state (imported unmodified from /root/reference, executed on CPU tensors with a numpy-backed `raymarching`
stand-in for its three integer ops, themselves pinned bit-exactly by tests/golden/raymarching.npz); generator:
tests/golden/make_golden_density_grid.py.  tests/test_oracle_cpu.py replays the fixture through this file.
"""
import ctypes

import numpy as np

from . import oracle as O

f32 = np.float32
INVALID = np.uint32(0xffffffff)


def _spread3(v):
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def morton3d(x, y, z):
    """raymarching.cu:56-72 (vectorised; checked against the scalar C oracle in tests)."""
    return _spread3(x) | (_spread3(y) << np.uint32(1)) | (_spread3(z) << np.uint32(2))


def _compact3(x):
    x = x.astype(np.uint32) & np.uint32(0x49249249)
    x = (x | (x >> np.uint32(2))) & np.uint32(0xc30c30c3)
    x = (x | (x >> np.uint32(4))) & np.uint32(0x0f00f00f)
    x = (x | (x >> np.uint32(8))) & np.uint32(0xff0000ff)
    x = (x | (x >> np.uint32(16))) & np.uint32(0x0000ffff)
    return x


def morton3d_invert(idx):
    """raymarching.cu:74-90 -> [N,3]."""
    idx = np.asarray(idx, dtype=np.uint32)
    return np.stack([_compact3(idx), _compact3(idx >> np.uint32(1)), _compact3(idx >> np.uint32(2))], axis=-1)


def cascade_scale(cas, bound, H):
    """renderer.py:416-419 / 472-474 / 503-505: python double arithmetic, rounded to fp32 when multiplied into a tensor."""
    bc = min(float(2 ** cas), float(f32(bound)))
    hg = bc / H
    return f32(bc - hg), f32(hg), f32(hg * 2.0)


CUDA_DIV = True   # tensor / scalar as torch's CUDA kernel does it (x * (1/s)); tests flip it to replay a CPU run of the reference


def cell_axis(c, H):
    """2 * coords.float() / (H - 1) - 1   (renderer.py:413,468,499)."""
    if CUDA_DIV:
        return (f32(2.0) * c.astype(f32)) * (f32(1.0) / f32(H - 1)) - f32(1.0)
    return (f32(2.0) * c.astype(f32)) / f32(H - 1) - f32(1.0)


def _jitter(coords, cas, bound, H, noise):
    s, hgs, _ = cascade_scale(cas, bound, H)
    p = cell_axis(coords, H) * s
    if noise is None:
        return p
    return p + (noise.astype(f32) * f32(2.0) - f32(1.0)) * hgs


def sample_full(C, H, bound, noise):
    """renderer.py:456-483.  noise [C,H^3,3] in meshgrid (x-major) order or None.  Returns xyzs [C,H^3,3] in Morton order."""
    H3 = H ** 3
    coords = morton3d_invert(np.arange(H3, dtype=np.uint32))          # cell of Morton index i
    lin = (coords[:, 0].astype(np.int64) * H + coords[:, 1]) * H + coords[:, 2]
    out = np.zeros((C, H3, 3), f32)
    for cas in range(C):
        out[cas] = _jitter(coords, cas, bound, H, None if noise is None else noise[cas][lin])
    return out


def occupied(grid):
    """torch.nonzero(density_grid[cas] > 0) (renderer.py:495): list of ascending index arrays."""
    return [np.nonzero(g > 0)[0].astype(np.uint32) for g in grid]


def sample_partial(C, H, bound, N, coords_rand, occ, noise, occ_pick_idx=None, occ_pick_u=None):
    """renderer.py:487-509.  Returns xyzs [C,2N,3], indices [C,2N] (uint32, INVALID where no sample exists)."""
    xyzs = np.zeros((C, 2 * N, 3), f32)
    indices = np.full((C, 2 * N), INVALID, np.uint32)
    for cas in range(C):
        cr = np.asarray(coords_rand[cas], dtype=np.uint32)
        idx_u = morton3d(cr[:, 0], cr[:, 1], cr[:, 2])
        nz = len(occ[cas])
        if nz > 0:
            if occ_pick_idx is not None:
                pick = np.asarray(occ_pick_idx[cas], dtype=np.int64)
            else:
                pick = np.minimum((np.asarray(occ_pick_u[cas], f32) * f32(nz)).astype(np.uint32), np.uint32(nz - 1)).astype(np.int64)
            idx_o = occ[cas][pick]
        else:
            idx_o = np.full(N, INVALID, np.uint32)
        idx = np.concatenate([idx_u, idx_o])
        coords = np.concatenate([cr, morton3d_invert(idx_o)], axis=0)
        p = _jitter(coords, cas, bound, H, None if noise is None else noise[cas])
        p[idx == INVALID] = 0
        xyzs[cas], indices[cas] = p, idx
    return xyzs, indices


def update(grid, indices, sigmas, density_scale, decay, density_thresh):
    """renderer.py:453,480-482,511-530.  grid [C,H^3]; indices [C,N] or None (identity); sigmas [C,N].
    Returns (new grid, mean_density fp32, threshold fp32, bitfield uint8).  Cells sampled more than once keep the largest
    sample (the reference's index_put keeps an arbitrary one of them)."""
    grid = np.array(grid, dtype=f32)
    C, H3 = grid.shape
    tmp = np.full_like(grid, -1.0)
    for cas in range(C):
        v = np.asarray(sigmas[cas], f32) * f32(density_scale) + f32(0.0)
        idx = np.arange(H3) if indices is None else np.asarray(indices[cas]).astype(np.int64)
        ok = (v >= 0) & (idx < H3)
        np.maximum.at(tmp[cas], idx[ok], v[ok])
    valid = (grid >= 0) & (tmp >= 0)
    grid[valid] = np.maximum(grid[valid] * f32(decay), tmp[valid])
    mean = f32(np.sum(np.clip(grid, 0, None), dtype=np.float64) / grid.size)
    thresh = min(mean, f32(density_thresh))
    return grid, mean, f32(thresh), O.packbits(grid, thresh)


def mark_untrained(poses, intrinsic, bound, C, H, grid):
    """renderer.py:380-442 -> (grid with -1 marks, count [C,H^3])."""
    poses = np.ascontiguousarray(poses, dtype=f32)
    grid = np.array(grid, dtype=f32).reshape(C, H ** 3)
    count = np.zeros((C, H ** 3), np.uint32)
    fx, fy, cx, cy = [float(v) for v in intrinsic]
    fn = O.lib().oracle_mark_untrained
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_uint32] + [ctypes.c_double] * 4 + [ctypes.c_float, ctypes.c_uint32, ctypes.c_uint32,
                                                                                ctypes.c_void_p, ctypes.c_void_p]
    fn(poses.ctypes.data, poses.shape[0], fx, fy, cx, cy, float(f32(bound)), C, H, grid.ctypes.data, count.ctypes.data)
    return grid, count
