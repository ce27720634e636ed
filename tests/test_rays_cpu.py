"""oracle/rays_oracle.py against the reference's own get_rays / srgb_to_linear outputs (tests/golden/rays.npz)."""
import os

import numpy as np
import pytest

from util import GOLDEN
from oracle import rays_oracle as RO

TOL = 1e-6


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "rays.npz"))


def test_get_rays_matches_reference(g):
    H, W = int(g["H"]), int(g["W"])
    o, d = RO.get_rays(g["poses"], g["intrinsics"], H, W)
    assert np.abs(d - g["all_d"]).max() <= TOL and (o == g["all_o"]).all()
    assert np.abs(np.linalg.norm(d, axis=-1) - 1).max() <= 2e-6
    o, d = RO.get_rays(g["poses"][:1], g["intrinsics"], H, W, g["rand_inds"])
    assert np.abs(d - g["rand_d"]).max() <= TOL and (o == g["rand_o"]).all()
    _, d = RO.get_rays(g["poses"][:1], g["intrinsics"], H, W, g["patch_inds"])
    assert np.abs(d - g["patch_d"]).max() <= TOL
    _, d = RO.get_rays(g["poses"][:2], g["intrinsics"], H, W, g["err_inds"])
    assert np.abs(d - g["err_d"]).max() <= TOL


def test_gather_pixels_matches_reference(g):
    px = RO.gather_pixels(g["images"], g["rand_inds"], image_index=[2])
    assert (px == g["px"]).all()
    gt = RO.gather_pixels(g["images"], g["rand_inds"], image_index=[2], gt=True, linear=True, bg=g["bg"])
    assert np.abs(gt - g["gt_linear_bg"]).max() <= TOL
    gt = RO.gather_pixels(g["images"], g["rand_inds"], image_index=[2], gt=True, linear=False, bg=1.0)
    assert np.abs(gt - g["gt_srgb_white"]).max() <= TOL
