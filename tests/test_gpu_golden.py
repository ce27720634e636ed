"""GPU: the CUDA path against the committed golden vectors (reference CUDA-extension outputs, tests/golden/)."""
import os

import numpy as np
import pytest
import torch

from util import GOLDEN, rel_err, canon_rays, gather_segments

pytestmark = pytest.mark.gpu


def _load(name):
    p = os.path.join(GOLDEN, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} missing")
    return np.load(p)


def test_grid_golden():
    from gridencoder.grid import grid_encode
    import _ngp_b200 as nb
    g = _load("grid.npz")
    od = torch.from_numpy(g["offsets"]).cuda(); x = torch.from_numpy(g["x"]).cuda()
    pls = float(g["per_level_scale"])
    for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
        emb = torch.from_numpy(g["table"]).to(dt).cuda()
        y = grid_encode(x, emb, od, pls, 16, False, 0, False, 0)
        np.testing.assert_array_equal(y.cpu().numpy(), g[f"y_{name}"])          # bit-exact features
        gr = torch.from_numpy(g["grad"]).to(dt).cuda()
        ge = torch.zeros_like(emb)
        nb.call("ngp_grid_encode_backward", gr.data_ptr(), x.data_ptr(), None, od.data_ptr(), ge.data_ptr(), 512, 3, 2, 8,
                float(np.log2(pls)), 16, None, None, 0, 0, 0, 1 if dt == torch.float16 else 0, 0)
        assert rel_err(ge.float().cpu().numpy(), g[f"grad_table_{name}"]) < (1e-5 if dt == torch.float32 else 2e-2)


def test_sh_golden():
    from shencoder import SHEncoder
    g = _load("sh.npz")
    y = SHEncoder(degree=8).cuda()(torch.from_numpy(g["dirs"]).cuda())
    assert np.abs(y.cpu().numpy() - g["y"]).max() < 2e-5 * max(1.0, np.abs(g["y"]).max())


def test_raymarching_golden():
    import raymarching
    import ngp_synth as S
    g = _load("raymarching.npz")
    grid = np.zeros(128 ** 3, np.float32); grid[g["occupied_cells"]] = 1.0
    bf = raymarching.packbits(torch.from_numpy(grid).view(1, -1).cuda(), 0.01)
    ro, rd = torch.from_numpy(g["rays_o"]).cuda(), torch.from_numpy(g["rays_d"]).cuda()
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0]).cuda()
    n, f = raymarching.near_far_from_aabb(ro, rd, aabb, 0.2)
    np.testing.assert_array_equal(n.cpu().numpy(), g["nears"]); np.testing.assert_array_equal(f.cpu().numpy(), g["fars"])
    import _ngp_b200 as nb
    N = ro.shape[0]; M = N * 512
    nz = torch.from_numpy(g["noises"]).cuda()
    for tag, dtg in (("g0", 0.0), ("g1", 1.0 / 128)):
        xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
        rays = torch.zeros(N, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
        nb.call("ngp_march_rays_train", ro.data_ptr(), rd.data_ptr(), bf.data_ptr(), 1.0, dtg, 1024, N, 1, 128, M, n.data_ptr(),
                f.data_ptr(), xyzs.data_ptr(), dirs.data_ptr(), deltas.data_ptr(), rays.data_ptr(), counter.data_ptr(), nz.data_ptr())
        assert counter.cpu().tolist() == g[f"counter_{tag}"].tolist()
        r = rays.cpu().numpy()
        np.testing.assert_array_equal(canon_rays(r)[:, [0, 2]], canon_rays(g[f"rays_{tag}"])[:, [0, 2]])
        np.testing.assert_array_equal(gather_segments(xyzs.cpu().numpy(), r), gather_segments(g[f"xyzs_{tag}"], g[f"rays_{tag}"]))
        np.testing.assert_array_equal(gather_segments(deltas.cpu().numpy(), r), gather_segments(g[f"deltas_{tag}"], g[f"rays_{tag}"]))
    # compositor on the golden sample layout
    rays_g = torch.from_numpy(g["rays_g0"]).cuda(); dl = torch.from_numpy(g["deltas_g0"]).cuda()
    ws, dp, im = raymarching.composite_rays_train(torch.from_numpy(g["sigmas"]).cuda(), torch.from_numpy(g["rgbs"]).cuda(), dl, rays_g, 1e-4)
    assert rel_err(ws.cpu().numpy(), g["weights_sum"]) < 1e-5 and rel_err(im.cpu().numpy(), g["image"]) < 1e-5 and rel_err(dp.cpu().numpy(), g["depth"]) < 1e-5


def test_ffmlp_golden():
    import _ngp_b200 as nb
    g = _load("ffmlp.npz")
    for tag, nl in (("sigma", 2), ("color", 3)):
        x = torch.from_numpy(g[f"{tag}_x"]).cuda(); w = torch.from_numpy(g[f"{tag}_w"]).cuda()
        fb = torch.empty(nl, 256, 64, dtype=torch.half, device="cuda"); y = torch.empty(256, 16, dtype=torch.half, device="cuda")
        nb.call("ngp_ffmlp_forward", x.data_ptr(), w.data_ptr(), 256, 32, 16, 64, nl, 0, 6, fb.data_ptr(), y.data_ptr())
        # reference = fp16-accumulate wmma; ours = fp32-accumulate tcgen05: within the reference's own rounding noise
        assert rel_err(y.cpu().numpy(), g[f"{tag}_y"]) < 3e-3 and rel_err(fb.cpu().numpy(), g[f"{tag}_fwd"]) < 3e-3
