"""CPU: pin the oracle (oracle/ngp_oracle.c + oracle/oracle.py) against the committed golden vectors, which
are outputs of the reference's own CUDA extensions recorded on a B200 (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from util import GOLDEN, rel_err, canon_rays, gather_segments
from oracle import oracle as O


def _load(name):
    p = os.path.join(GOLDEN, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not generated yet")
    return np.load(p)


def test_offsets_table_matches_survey():
    offsets, pls = O.grid_offsets(3, 16, 2, 2, 16, 19, 2048)
    assert np.diff(offsets)[:6].tolist() == [4920, 13824, 32768, 85184, 216000, 524288]
    assert int(offsets[-1]) == 6119864 and abs(pls - 2 ** (7 / 15)) < 1e-12


def test_grid_forward_golden():
    g = _load("grid.npz")
    S = float(np.log2(g["per_level_scale"]))
    for name, dt in (("f32", np.float32), ("f16", np.float16)):
        y, idx, dy = O.grid_forward(g["x"], g["table"].astype(dt), g["offsets"], S, 16, scales=g["level_scales"],
                                    want_indices=True, want_dy_dx=True)
        # with the device's level scales the restatement reproduces the reference kernel bit-for-bit
        np.testing.assert_array_equal(y.view(np.uint8), g[f"y_{name}"].view(np.uint8))
        np.testing.assert_array_equal(dy.view(np.uint8), g[f"dy_dx_{name}"].view(np.uint8))
        # with libm's exp2f the level scales differ from the GPU's MUFU.EX2 by <= 1 ulp: same cells, last-bit
        # differences in the interpolation weights
        y2 = O.grid_forward(g["x"], g["table"].astype(dt), g["offsets"], S, 16)
        assert rel_err(y2, g[f"y_{name}"]) < (1e-4 if dt == np.float32 else 2e-3)
    assert np.all(idx[16:] < np.diff(g["offsets"])[None, :, None])


def test_grid_backward_golden():
    g = _load("grid.npz")
    S = float(np.log2(g["per_level_scale"]))
    n = int(g["offsets"][-1])
    for name, dt, tol in (("f32", np.float32, 1e-5), ("f16", np.float16, 2e-2)):
        gt = O.grid_backward(g["grad"].astype(dt), g["x"], g["offsets"], n, 2, S, 16, scales=g["level_scales"])
        assert rel_err(gt, g[f"grad_table_{name}"]) < tol
        assert not np.any(g[f"grad_table_{name}"][gt == 0] != 0)


def test_grid_properties():
    """Linearity in the table and adjointness of forward/backward (fp32), on random data."""
    rng = np.random.default_rng(0)
    offsets, pls = O.grid_offsets(3, 6, 2, 2, 16, 10, 128)
    S = float(np.log2(pls))
    x = rng.random((300, 3), dtype=np.float32)
    t1 = rng.standard_normal((offsets[-1], 2)).astype(np.float32)
    t2 = rng.standard_normal((offsets[-1], 2)).astype(np.float32)
    y1, y2 = O.grid_forward(x, t1, offsets, S, 16), O.grid_forward(x, t2, offsets, S, 16)
    y12 = O.grid_forward(x, t1 + 2 * t2, offsets, S, 16)
    assert rel_err(y1 + 2 * y2, y12) < 1e-5
    g = rng.standard_normal(y1.shape).astype(np.float32)
    gt = O.grid_backward(g, x, offsets, int(offsets[-1]), 2, S, 16)
    assert abs((y1.astype(np.float64) * g).sum() - (t1 * gt).sum()) < 1e-3 * abs((t1 * gt).sum())
    # out-of-range points produce zeros and no gradient
    xo = x.copy(); xo[:10, 0] = 1.5
    assert np.all(O.grid_forward(xo, t1, offsets, S, 16)[:10] == 0)


def test_sh_golden():
    g = _load("sh.npz")
    y = O.sh_encode(g["dirs"][:128], 8)          # unit vectors only (scipy oracle domain)
    assert np.abs(y - g["y"][:128]).max() < 2e-5


def _bitfield(g):
    grid = np.zeros(128 ** 3, np.float32)
    grid[g["occupied_cells"]] = 1.0
    return O.packbits(grid, 0.01)


def test_raymarching_golden():
    g = _load("raymarching.npz")
    bf = _bitfield(g)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    n, f = O.near_far_from_aabb(g["rays_o"], g["rays_d"], aabb, 0.2)
    np.testing.assert_array_equal(n, g["nears"]); np.testing.assert_array_equal(f, g["fars"])
    N = len(n)
    for tag, dtg in (("g0", 0.0), ("g1", 1.0 / 128)):
        x, d, l, r, c = O.march_rays_train(g["rays_o"], g["rays_d"], bf, 1.0, dtg, 1024, 1, 128, N * 512, n, f, g["noises"])
        assert c.tolist() == g[f"counter_{tag}"].tolist()
        np.testing.assert_array_equal(canon_rays(r)[:, [0, 2]], canon_rays(g[f"rays_{tag}"])[:, [0, 2]])   # bit-exact counts
        np.testing.assert_array_equal(gather_segments(x, r), gather_segments(g[f"xyzs_{tag}"], g[f"rays_{tag}"]))
        np.testing.assert_array_equal(gather_segments(l, r), gather_segments(g[f"deltas_{tag}"], g[f"rays_{tag}"]))


def test_composite_golden():
    g = _load("raymarching.npz")
    rays, dl = g["rays_g0"], g["deltas_g0"]
    ws, dp, im = O.composite_rays_train_forward(g["sigmas"], g["rgbs"], dl, rays, 1e-4)
    # expf (CPU) vs __expf (device): 1e-5 of the output scale
    assert rel_err(ws, g["weights_sum"]) < 1e-5 and rel_err(dp, g["depth"]) < 1e-5 and rel_err(im, g["image"]) < 1e-5
    gs, gc = O.composite_rays_train_backward(g["grad_ws"], g["grad_image"], g["sigmas"], g["rgbs"], dl, rays, ws, im, 1e-4)
    assert rel_err(gs, g["grad_sigmas"]) < 1e-4 and rel_err(gc, g["grad_rgbs"]) < 1e-5


def test_inference_golden():
    g = _load("raymarching.npz")
    bf = _bitfield(g)
    N = len(g["nears"])
    alive = np.arange(N, dtype=np.int32)
    x, d, l = O.march_rays(N, 4, alive, g["nears"], g["rays_o"], g["rays_d"], 1.0, 0.0, 1024, 1, 128, bf, g["nears"],
                           g["fars"], np.zeros(N, np.float32), align=128)
    np.testing.assert_array_equal(x, g["inf_xyzs"]); np.testing.assert_array_equal(l, g["inf_deltas"])
    a, t, ws, dp, im = O.composite_rays(N, 4, 1e-2, alive, g["nears"], g["inf_sigmas"], g["inf_rgbs"], l,
                                        np.zeros(N), np.zeros(N), np.zeros((N, 3)))
    np.testing.assert_array_equal(a, g["inf_alive"])
    assert rel_err(t, g["inf_t"]) < 1e-6 and rel_err(ws, g["inf_ws"]) < 1e-5 and rel_err(im, g["inf_image"]) < 1e-5


def test_mlp_golden():
    """numpy MLP (fp32 accumulate) vs the reference's wmma/CUTLASS kernels (fp16 accumulate): the reference's own
    accumulation error bounds the difference (3e-3 of the output scale; weight grads 2e-2, fp16 split-K)."""
    g = _load("ffmlp.npz")
    for tag, nl in (("sigma", 2), ("color", 3)):
        y, fwd = O.mlp_forward(g[f"{tag}_x"], g[f"{tag}_w"], 32, 64, nl)
        assert rel_err(y, g[f"{tag}_y"]) < 3e-3 and rel_err(fwd, g[f"{tag}_fwd"]) < 3e-3
        np.testing.assert_array_equal(g[f"{tag}_y"], g[f"{tag}_y_inf"])
        gi, gw, bwd = O.mlp_backward(g[f"{tag}_g"], g[f"{tag}_x"], g[f"{tag}_w"], g[f"{tag}_fwd"], 32, 64, nl)
        assert rel_err(bwd, g[f"{tag}_bwd"]) < 5e-3 and rel_err(gi, g[f"{tag}_gi"]) < 5e-3
        assert rel_err(gw, g[f"{tag}_gw"].astype(np.float32)) < 2e-2


def test_morton_roundtrip_and_packbits():
    rng = np.random.default_rng(1)
    c = rng.integers(0, 1024, (500, 3))
    idx = O.morton3D(c)
    np.testing.assert_array_equal(O.morton3D_invert(idx), c)
    import ngp_synth as S
    np.testing.assert_array_equal(S.morton3d_np(c[:, 0], c[:, 1], c[:, 2]).astype(np.int32), idx)
    grid = rng.random(4096).astype(np.float32)
    np.testing.assert_array_equal(O.packbits(grid, 0.5), S.packbits_np(grid, 0.5))
