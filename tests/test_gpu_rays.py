"""ngp_get_rays / ngp_gather_pixels against the oracle, the reference-generated fixture and the reference's
RNG consumption.  Floating-point tolerance 1e-6 (torch.norm / cuBLAS accumulation order is not specified, see rays.cu)."""
import os

import numpy as np
import pytest
import torch

from util import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-6


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "rays.npz"))


def test_fixture(g):
    import ngp_rays
    from oracle import rays_oracle as RO
    H, W = int(g["H"]), int(g["W"])
    poses = torch.from_numpy(g["poses"]).cuda()
    o, d = ngp_rays.rays_from_pixels(poses, g["intrinsics"], H, W)
    assert np.abs(d.cpu().numpy() - g["all_d"]).max() <= TOL and (o.cpu().numpy() == g["all_o"]).all()
    for key, nb in (("rand", 1), ("patch", 1), ("err", 2)):
        inds = torch.from_numpy(g[f"{key}_inds"]).cuda()
        o, d = ngp_rays.rays_from_pixels(poses[:nb], g["intrinsics"], H, W, inds)
        assert np.abs(d.cpu().numpy() - g[f"{key}_d"]).max() <= TOL
        oo, od = RO.get_rays(g["poses"][:nb], g["intrinsics"], H, W, g[f"{key}_inds"])
        assert np.abs(d.cpu().numpy() - od).max() <= TOL and (o.cpu().numpy() == oo).all()
    images = torch.from_numpy(g["images"]).cuda()
    inds = torch.from_numpy(g["rand_inds"]).cuda()
    px = ngp_rays.gather_pixels(images, inds, image_index=[2])
    assert (px.cpu().numpy() == g["px"]).all()
    gt = ngp_rays.gather_pixels(images, inds, image_index=[2], gt=True, linear=True, bg_color=torch.from_numpy(g["bg"]).cuda())
    assert np.abs(gt.cpu().numpy() - g["gt_linear_bg"]).max() <= TOL
    gt = ngp_rays.gather_pixels(images, inds, image_index=[2], gt=True, linear=False, bg_color=1.0)
    assert np.abs(gt.cpu().numpy() - g["gt_srgb_white"]).max() <= TOL


def test_full_frame_and_uint8():
    """800x800 frame, 4 cameras, against the oracle; uint8 RGB images (C == 3: no blend)."""
    import ngp_rays
    import ngp_synth as S
    from oracle import rays_oracle as RO
    poses = S.make_cameras(4, seed=11).float()
    intr = S.intrinsics()
    o, d = ngp_rays.rays_from_pixels(poses.cuda(), intr, 800, 800)
    oo, od = RO.get_rays(poses.numpy(), intr, 800, 800)
    assert np.abs(d.cpu().numpy() - od).max() <= TOL and (o.cpu().numpy() == oo).all()
    assert np.abs(np.linalg.norm(d.cpu().numpy(), axis=-1) - 1).max() <= 2e-6
    # agrees with the synthetic-scene helper the bench uses
    ro, rd = S.get_rays(poses[1], intr, 800, 800, device="cuda")
    assert (rd - d[1]).abs().max().item() <= 2e-6
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (2, 32, 48, 3), dtype=np.uint8)
    inds = rng.integers(0, 32 * 48, (2, 100))
    out = ngp_rays.gather_pixels(torch.from_numpy(img).cuda(), torch.from_numpy(inds).cuda(), gt=True)
    assert (out.cpu().numpy() == RO.gather_pixels(img, inds, gt=True)).all()
    with pytest.raises(RuntimeError, match="3 or 4 channels"):
        ngp_rays.gather_pixels(torch.zeros(1, 4, 4, 2).cuda(), torch.zeros(1, 3, dtype=torch.long).cuda())


def test_get_rays_consumes_rng_like_reference(g):
    """same seed -> same pixel choice as the reference's get_rays for the three sampling modes (the fixture holds the reference's
    indices; CUDA and CPU generators differ, so the comparison replays the reference's draws on the CPU generator's device)."""
    import ngp_rays
    H, W = int(g["H"]), int(g["W"])
    poses = torch.from_numpy(g["poses"]).cuda()
    torch.manual_seed(7)
    r = ngp_rays.get_rays(poses[:1], g["intrinsics"], H, W, 50)
    torch.manual_seed(7)
    expect = torch.randint(0, H * W, size=[50], device="cuda")
    assert (r["inds"][0] == expect).all() and tuple(r["rays_d"].shape) == (1, 50, 3)
    torch.manual_seed(8)
    r = ngp_rays.get_rays(poses[:1], g["intrinsics"], H, W, 64, patch_size=4)
    inds = r["inds"][0].view(4, 4, 4)
    assert (inds[:, 1:, :] - inds[:, :-1, :] == W).all() and (inds[:, :, 1:] - inds[:, :, :-1] == 1).all()
    torch.manual_seed(9)
    r = ngp_rays.get_rays(poses[:2], g["intrinsics"], H, W, 40, error_map=torch.from_numpy(g["err_map"]))
    assert tuple(r["inds"].shape) == (2, 40) and tuple(r["inds_coarse"].shape) == (2, 40)
    rows, cols = r["inds"] // W, r["inds"] % W
    assert ((rows == (r["inds_coarse"] // 128 * (H / 128)).long()) | (rows == (r["inds_coarse"] // 128 * (H / 128)).long() + 1)).all()
    r = ngp_rays.get_rays(poses, g["intrinsics"], H, W, -1)
    assert "inds" not in r and np.abs(r["rays_d"].cpu().numpy() - g["all_d"]).max() <= TOL
