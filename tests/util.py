"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def rel_err(a, b):
    """max |a-b| / max|b| — norm-wise relative error (fp16 values cannot meet 1e-3 element-wise: 1 ulp = 9.8e-4)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def canon_rays(rays):
    """Sort the (ray id, offset, count) table by ray id: the reference's row order is atomics-dependent."""
    rays = np.asarray(rays)
    return rays[np.argsort(rays[:, 0], kind="stable")]


def gather_segments(buf, rays):
    """Concatenate per-ray sample segments in ray-id order -> layout-independent view of a marcher output."""
    buf = np.asarray(buf); r = canon_rays(rays)
    return np.concatenate([buf[o:o + c] for _, o, c in r if c > 0] or [buf[:0]])


def synth_rays(N, seed=3, device="cpu"):
    """N rays of a synthetic 800x800 camera looking at the unit cube + a box-union bitfield (numpy/torch CPU)."""
    import ngp_synth as S
    poses = S.make_cameras(4, seed=11)
    intr = S.intrinsics()
    inds = torch.randint(0, 800 * 800, (N,), generator=gen(seed))
    rays_o, rays_d = S.get_rays(poses[seed % 4], intr, 800, 800, inds)
    grid, fill = S.box_union_density(128, seed=12)
    bitfield = torch.from_numpy(S.packbits_np(grid.numpy()))
    return rays_o.to(device), rays_d.to(device), bitfield.to(device), grid


# ---- occupancy-grid fixtures -------------------------------------------------------------------
def lex_to_morton(H):
    """Morton index of the cells in the reference's meshgrid (x-major) enumeration."""
    from oracle.density_grid_oracle import morton3d
    xs, ys, zs = [a.reshape(-1).astype(np.uint32) for a in np.meshgrid(np.arange(H), np.arange(H), np.arange(H), indexing="ij")]
    return morton3d(xs, ys, zs)


def check_duplicates_aware(new, ref, before, indices, sigmas, decay):
    """cells sampled once must agree exactly; for cells drawn several times the reference keeps an arbitrary sample, the oracle
    the largest: the reference value must be the EMA of ONE of the candidates."""
    C, H3 = new.shape
    for cas in range(C):
        diff = np.nonzero(new[cas] != ref[cas])[0]
        for cell in diff:
            cand = np.asarray(sigmas[cas], np.float32)[np.asarray(indices[cas]) == cell]
            assert len(cand) > 1, f"cell {cell} sampled once but differs"
            allowed = np.maximum(before[cas, cell] * np.float32(decay), cand)
            assert ref[cas, cell] in allowed
            assert new[cas, cell] == allowed.max()
    return sum(int((new[c] != ref[c]).sum()) for c in range(C))
