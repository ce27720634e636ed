"""Oracle of the occupancy-grid maintenance (oracle/density_grid_oracle.py) against the reference's own
mark_untrained_grid / update_extra_state outputs (tests/golden/density_grid.npz, produced by running
/root/reference/nerf/renderer.py unmodified on CPU — tests/golden/make_golden_density_grid.py)."""
import os

import numpy as np
import pytest

from util import GOLDEN, lex_to_morton, check_duplicates_aware
from oracle import oracle as O
from oracle import density_grid_oracle as DG


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "density_grid.npz"))


def test_vector_morton_matches_scalar_oracle():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 1024, (200, 3)).astype(np.uint32)
    v = DG.morton3d(c[:, 0], c[:, 1], c[:, 2])
    assert (v.astype(np.int64) == O.morton3D(c).astype(np.int64) % (1 << 32)).all()
    assert (DG.morton3d_invert(v) == c).all()


def test_mark_untrained_matches_reference(g):
    H, C, bound = int(g["H"]), int(g["C"]), float(g["bound"])
    grid, count = DG.mark_untrained(g["poses"], g["intrinsic"], bound, C, H, np.zeros((C, H ** 3), np.float32))
    assert ((grid < 0) == (count == 0)).all()
    assert (grid == g["marked_grid"]).all()           # exact at this size (cuBLAS/FMA order may flip boundary cells in general)
    assert 0 < (grid < 0).sum() < grid.size


@pytest.mark.parametrize("cuda_div", [False, True])
def test_update_extra_state_replay(g, cuda_div):
    """Replays the four reference updates.  With CUDA_DIV off (true division, as the CPU run of the reference computed) the
    sample positions are bit-identical; with the GPU rule (x * (1/15)) they differ by at most one ulp."""
    H, C, bound = int(g["H"]), int(g["C"]), float(g["bound"])
    H3, N = H ** 3, H ** 3 // 4
    lm = lex_to_morton(H)
    DG.CUDA_DIV = cuda_div
    try:
        grid = g["marked_grid"].copy()
        for it in range(4):
            ref_xyz, ref_sig = g[f"u{it}_xyzs"], g[f"u{it}_sigmas"]
            if bool(g[f"u{it}_full"]):
                x = DG.sample_full(C, H, bound, g[f"u{it}_noise"])[:, lm, :]
                sig = np.zeros((C, H3), np.float32); sig[:, lm] = ref_sig
                idx = None
            else:
                x, idx = DG.sample_partial(C, H, bound, N, g[f"u{it}_coords"], DG.occupied(grid), g[f"u{it}_noise"],
                                           occ_pick_idx=g[f"u{it}_picks"])
                sig = ref_sig
            if cuda_div:
                assert np.abs(x - ref_xyz).max() <= 2.4e-7
            else:
                assert (x == ref_xyz).all()
            new, mean, thresh, bits = DG.update(grid, idx, sig, float(g["density_scale"]), float(g["decay"]), float(g["density_thresh"]))
            ref_grid = g[f"u{it}_grid"]
            if idx is None:
                assert (new == ref_grid).all()
                assert abs(float(mean) - float(g[f"u{it}_mean"])) <= 1e-6 * float(mean)
                assert (bits == g[f"u{it}_bitfield"]).all()
            else:
                ndiff = check_duplicates_aware(new, ref_grid, grid, idx, sig, float(g["decay"]))
                assert ndiff < 0.1 * new.size
                assert (O.packbits(ref_grid, min(float(g[f"u{it}_mean"]), float(g["density_thresh"]))) == g[f"u{it}_bitfield"]).all()
            grid = ref_grid.copy()
    finally:
        DG.CUDA_DIV = True


def test_update_properties():
    """size-independent properties: untouched cells keep their value, -1 cells stay -1, bitfield == packbits(grid, thresh),
    updating twice with the same samples is the EMA recurrence."""
    rng = np.random.default_rng(3)
    C, H = 2, 8
    H3 = H ** 3
    grid = rng.random((C, H3)).astype(np.float32)
    grid[rng.random((C, H3)) < 0.2] = -1
    idx = rng.integers(0, H3, (C, 100)).astype(np.uint32)
    sig = rng.random((C, 100)).astype(np.float32) * 3
    new, mean, thresh, bits = DG.update(grid, idx, sig, 2.0, 0.9, 0.5)
    for cas in range(C):
        untouched = np.ones(H3, bool); untouched[idx[cas]] = False
        assert (new[cas][untouched] == grid[cas][untouched]).all()
    assert ((new < 0) == (grid < 0)).all()
    assert (bits == O.packbits(new, thresh)).all()
    assert thresh == min(mean, np.float32(0.5))
    assert (new[grid >= 0] >= grid[grid >= 0] * np.float32(0.9)).all()
