"""Occupancy-grid maintenance kernels (ngp_density_grid_*) against the oracle
(oracle/density_grid_oracle.py), the reference-generated fixture (tests/golden/density_grid.npz) and — seed for seed —
the reference's op-by-op torch sequence restated in nerf_step.NeRFFieldFF.update_extra_state_unfused."""
import os

import numpy as np
import pytest
import torch

from util import GOLDEN, gen, lex_to_morton, check_duplicates_aware

pytestmark = pytest.mark.gpu


def _nb():
    import _ngp_b200 as nb
    return nb


def _dev(a, dt=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dt) if dt is not None else t).cuda()


def k_mark(poses, intr, bound, C, H, grid):
    nb = _nb()
    g = _dev(grid, torch.float32).view(C, H ** 3).clone()
    count = torch.empty(C, H ** 3, dtype=torch.int32, device="cuda")
    nm = torch.zeros(1, dtype=torch.int32, device="cuda")
    p = _dev(poses, torch.float32)
    nb.call("ngp_density_grid_mark_untrained", p.data_ptr(), p.shape[0], *[float(v) for v in intr], float(bound), C, H,
            g.data_ptr(), count.data_ptr(), nm.data_ptr())
    return g.cpu().numpy(), count.cpu().numpy().astype(np.uint32), int(nm.item())


def k_occupied(grid, C, H):
    nb = _nb(); lib = nb.load()
    g = _dev(grid, torch.float32)
    occ = torch.full((C, H ** 3), -1, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(C, dtype=torch.int32, device="cuda")
    scratch = torch.empty(max(1, lib.ngp_density_grid_occupied_scratch_bytes(C, H) // 4), dtype=torch.int32, device="cuda")
    nb.call("ngp_density_grid_occupied", g.data_ptr(), C, H, occ.data_ptr(), cnt.data_ptr(), scratch.data_ptr())
    return occ, cnt


def k_sample_full(C, H, bound, noise):
    nb = _nb()
    x = torch.empty(C, H ** 3, 3, device="cuda")
    nz = None if noise is None else _dev(noise, torch.float32)
    nb.call("ngp_density_grid_sample_full", C, H, float(bound), nb.ptr(nz), x.data_ptr())
    return x.cpu().numpy()


def k_sample_partial(C, H, bound, N, coords, occ, cnt, noise, pick_idx=None, pick_u=None):
    nb = _nb()
    x = torch.empty(C, 2 * N, 3, device="cuda")
    idx = torch.empty(C, 2 * N, dtype=torch.int32, device="cuda")
    c = _dev(coords, torch.int32)
    pi = None if pick_idx is None else _dev(pick_idx, torch.int64)
    pu = None if pick_u is None else _dev(pick_u, torch.float32)
    nz = _dev(noise, torch.float32)
    nb.call("ngp_density_grid_sample_partial", C, H, float(bound), N, c.data_ptr(), nb.ptr(pi), nb.ptr(pu), occ.data_ptr(),
            cnt.data_ptr(), nz.data_ptr(), x.data_ptr(), idx.data_ptr())
    return x.cpu().numpy(), idx.cpu().numpy().astype(np.uint32)


def k_update(grid, indices, sigmas, N, density_scale, decay, density_thresh, C, H):
    nb = _nb(); lib = nb.load()
    g = _dev(grid, torch.float32).view(C, H ** 3).clone()
    tmp = torch.empty_like(g)
    idx = None if indices is None else _dev(indices.astype(np.int64), torch.int64).to(torch.int32)
    s = _dev(sigmas, torch.float32)
    bits = torch.zeros(C * H ** 3 // 8, dtype=torch.uint8, device="cuda")
    state = torch.zeros(2, device="cuda")
    scratch = torch.empty(max(1, lib.ngp_density_grid_update_scratch_bytes(C, H) // 8), dtype=torch.float64, device="cuda")
    nb.call("ngp_density_grid_update", g.data_ptr(), tmp.data_ptr(), nb.ptr(idx), s.data_ptr(), N, float(density_scale), float(decay),
            float(density_thresh), C, H, bits.data_ptr(), state.data_ptr(), scratch.data_ptr())
    st = state.cpu().numpy()
    return g.cpu().numpy(), st[0], st[1], bits.cpu().numpy()


def test_golden_fixture_replay():
    """the four reference updates of the fixture through the kernels: positions within 1 ulp of the reference's CPU run (and
    bit-identical to the oracle's GPU-division rule), grids / bitfields as the oracle test establishes for the oracle."""
    from oracle import density_grid_oracle as DG
    from oracle import oracle as O
    g = np.load(os.path.join(GOLDEN, "density_grid.npz"))
    H, C, bound = int(g["H"]), int(g["C"]), float(g["bound"])
    H3, N = H ** 3, H ** 3 // 4
    lm = lex_to_morton(H).astype(np.int64)
    # mark_untrained
    grid, count, nm = k_mark(g["poses"], g["intrinsic"], bound, C, H, np.zeros((C, H3), np.float32))
    og, oc = DG.mark_untrained(g["poses"], g["intrinsic"], bound, C, H, np.zeros((C, H3), np.float32))
    assert (count == oc).all() and (grid == og).all() and nm == int((og < 0).sum())
    assert (grid == g["marked_grid"]).all()
    grid = g["marked_grid"].copy()
    for it in range(4):
        ref_xyz, ref_sig, ref_grid = g[f"u{it}_xyzs"], g[f"u{it}_sigmas"], g[f"u{it}_grid"]
        if bool(g[f"u{it}_full"]):
            x = k_sample_full(C, H, bound, g[f"u{it}_noise"])
            assert (x == DG.sample_full(C, H, bound, g[f"u{it}_noise"])).all()
            assert np.abs(x[:, lm, :] - ref_xyz).max() <= 2.4e-7
            sig = np.zeros((C, H3), np.float32); sig[:, lm] = ref_sig
            new, mean, thresh, bits = k_update(grid, None, sig, H3, 1.0, float(g["decay"]), float(g["density_thresh"]), C, H)
            assert (new == ref_grid).all()
            assert abs(float(mean) - float(g[f"u{it}_mean"])) <= 1e-6 * float(mean)
            assert (bits == g[f"u{it}_bitfield"]).all()
        else:
            occ, cnt = k_occupied(grid, C, H)
            oo = DG.occupied(grid)
            for cas in range(C):
                assert int(cnt[cas]) == len(oo[cas])
                assert (occ[cas, :len(oo[cas])].cpu().numpy().astype(np.uint32) == oo[cas]).all()
            x, idx = k_sample_partial(C, H, bound, N, g[f"u{it}_coords"], occ, cnt, g[f"u{it}_noise"], pick_idx=g[f"u{it}_picks"])
            ox, oidx = DG.sample_partial(C, H, bound, N, g[f"u{it}_coords"], oo, g[f"u{it}_noise"], occ_pick_idx=g[f"u{it}_picks"])
            assert (idx == oidx).all() and (x == ox).all()
            assert np.abs(x - ref_xyz).max() <= 2.4e-7
            new, mean, thresh, bits = k_update(grid, idx, ref_sig, 2 * N, 1.0, float(g["decay"]), float(g["density_thresh"]), C, H)
            onew, omean, othresh, obits = DG.update(grid, idx, ref_sig, 1.0, float(g["decay"]), float(g["density_thresh"]))
            assert (new == onew).all() and (bits == O.packbits(new, thresh)).all()
            assert abs(float(mean) - float(omean)) <= 1e-6 * float(omean)
            check_duplicates_aware(new, ref_grid, grid, idx, ref_sig, float(g["decay"]))
        grid = ref_grid.copy()


@pytest.mark.parametrize("C,H,bound", [(1, 128, 1.0), (3, 32, 4.0), (2, 64, 1.5), (1, 4, 1.0)])
def test_kernels_vs_oracle_random(C, H, bound):
    """seeded random grids at several sizes incl. the full 128^3: every stage bit-exact against the oracle."""
    from oracle import density_grid_oracle as DG
    from oracle import oracle as O
    rng = np.random.default_rng(C * 1000 + H)
    H3, N = H ** 3, max(1, H ** 3 // 4)
    grid = (rng.random((C, H3)).astype(np.float32) ** 8) * 5
    grid[rng.random((C, H3)) < 0.3] = 0
    grid[rng.random((C, H3)) < 0.1] = -1
    # occupied list
    occ, cnt = k_occupied(grid, C, H)
    oo = DG.occupied(grid)
    for cas in range(C):
        assert int(cnt[cas]) == len(oo[cas])
        assert (occ[cas, :len(oo[cas])].cpu().numpy().astype(np.uint32) == oo[cas]).all()
    # full sampling, with and without jitter
    noise = rng.random((C, H3, 3)).astype(np.float32)
    assert (k_sample_full(C, H, bound, noise) == DG.sample_full(C, H, bound, noise)).all()
    assert (k_sample_full(C, H, bound, None) == DG.sample_full(C, H, bound, None)).all()
    # partial sampling through the device-side pick
    coords = rng.integers(0, H, (C, N, 3)).astype(np.int32)
    pu = rng.random((C, N)).astype(np.float32)
    pn = rng.random((C, 2 * N, 3)).astype(np.float32)
    x, idx = k_sample_partial(C, H, bound, N, coords, occ, cnt, pn, pick_u=pu)
    ox, oidx = DG.sample_partial(C, H, bound, N, coords, oo, pn, occ_pick_u=pu)
    assert (idx == oidx).all() and (x == ox).all()
    lim = bound * 1.0000001
    assert np.abs(x).max() <= lim
    # update (duplicates present), scale != 1
    sig = (rng.random((C, 2 * N)).astype(np.float32) ** 4) * 10
    new, mean, thresh, bits = k_update(grid, idx, sig, 2 * N, 1.7, 0.95, 0.01, C, H)
    onew, omean, othresh, obits = DG.update(grid, idx, sig, 1.7, 0.95, 0.01)
    assert (new == onew).all()
    assert abs(float(mean) - float(omean)) <= 1e-6 * abs(float(omean))
    assert thresh == min(mean, np.float32(0.01))
    assert (bits == O.packbits(new, thresh)).all()
    # full update: identity scatter
    sigf = rng.random((C, H3)).astype(np.float32)
    new2, mean2, thresh2, bits2 = k_update(new, None, sigf, H3, 1.0, 0.5, 100.0, C, H)
    onew2, omean2, _, _ = DG.update(new, None, sigf, 1.0, 0.5, 100.0)
    assert (new2 == onew2).all() and thresh2 == mean2
    assert (bits2 == O.packbits(new2, thresh2)).all()


def test_empty_occupied_list_and_bad_inputs():
    """no occupied cell: the second half of a partial update yields no samples (indices 0xffffffff) and the update ignores
    them; out-of-range coordinates likewise; invalid sizes are refused with an error message."""
    nb = _nb()
    C, H, N = 1, 8, 16
    grid = np.zeros((C, H ** 3), np.float32)
    occ, cnt = k_occupied(grid, C, H)
    assert int(cnt[0]) == 0
    coords = np.zeros((C, N, 3), np.int32); coords[0, 0] = (H, 0, 0)
    x, idx = k_sample_partial(C, H, 1.0, N, coords, occ, cnt, np.zeros((C, 2 * N, 3), np.float32), pick_u=np.zeros((C, N), np.float32))
    assert (idx[0, N:] == 0xffffffff).all() and idx[0, 0] == 0xffffffff and (idx[0, 1:N] == 0).all()
    new, mean, thresh, bits = k_update(grid, idx, np.ones((C, 2 * N), np.float32), 2 * N, 1.0, 0.95, 0.01, C, H)
    assert new[0, 0] == 1.0 and (new[0, 1:] == 0).all()
    with pytest.raises(RuntimeError, match="power of two"):
        k_sample_full(1, 12, 1.0, None)
    with pytest.raises(RuntimeError, match="cascade"):
        k_sample_full(0, 8, 1.0, None)


def test_mark_untrained_full_size():
    """128^3 x 2 cascades, 100 cameras: counts bit-exact against the C oracle."""
    from oracle import density_grid_oracle as DG
    import ngp_synth as S
    poses = S.make_cameras(100, seed=11).numpy()
    intr = S.intrinsics()
    grid = np.zeros((2, 128 ** 3), np.float32)
    g, count, nm = k_mark(poses, intr, 2.0, 2, 128, grid)
    og, oc = DG.mark_untrained(poses, intr, 2.0, 2, 128, grid)
    assert (count == oc).all() and (g == og).all()
    assert 0 < nm < g.size


def _field(bound=1, H=128):
    import ngp_synth as S
    from nerf_step import NeRFFieldFF
    torch.manual_seed(0)
    m = NeRFFieldFF(bound=bound, grid_size=H, fused=True).cuda()
    with torch.no_grad():
        m.encoder.embeddings.uniform_(-1.0, 1.0)       # spread sigma = exp(h) over the threshold
    return m


@pytest.mark.parametrize("bound,H", [(1, 128), (2, 64)])
def test_update_extra_state_vs_reference_sequence(bound, H):
    """Seed for seed against the reference's torch-op sequence (restated in update_extra_state_unfused, density through the
    module-by-module GridEncoder -> FFMLP -> trunc_exp path): 2 full + 2 partial updates.  Full updates draw identical
    random numbers, so the grids agree exactly; the fused density kernel computes the same numbers as the module path."""
    a, b = _field(bound, H), _field(bound, H)
    b.load_state_dict(a.state_dict())
    for it in range(4):
        if it == 2:
            a.iter_density = b.iter_density = 16
        full = a.iter_density < 16
        torch.manual_seed(100 + it)
        with torch.autocast("cuda", dtype=torch.float16):
            a.update_extra_state_unfused()
        torch.manual_seed(100 + it)
        b.update_extra_state(exact_rng=True)
        ga, gb = a.density_grid.cpu().numpy(), b.density_grid.cpu().numpy()
        if full:
            # identical samples; the fused kernel's expf and torch.exp may differ in the last bit
            assert np.abs(ga - gb).max() <= 2.5e-7 * np.abs(ga).max() and (ga == gb).mean() > 0.99
            assert int((a.density_bitfield != b.density_bitfield).sum()) <= 2
            b.density_grid.copy_(a.density_grid); b.density_bitfield.copy_(a.density_bitfield)
        else:
            # same draws, but cells drawn twice keep an arbitrary sample in torch's index_put and the largest here
            frac = float((ga != gb).mean())
            assert frac < 0.08
            same = ga == gb
            assert (gb[~same] >= ga[~same]).all()
            b.density_grid.copy_(a.density_grid); b.density_bitfield.copy_(a.density_bitfield)
        assert abs(a.mean_density - b.mean_density) <= 2e-6 * max(abs(a.mean_density), 1e-6) or not full
        assert a.iter_density == b.iter_density


def test_update_extra_state_device_pick_and_marcher():
    """default (host-sync-free pick) partial updates keep the grid consistent: bitfield == packbits(grid, min(mean, thresh)),
    -1 cells stay -1, and the marcher consumes the refreshed bitfield."""
    import raymarching
    import ngp_synth as S
    from oracle import oracle as O
    m = _field(1, 128)
    m.mark_untrained_grid(S.make_cameras(3, radius=1.5, seed=5), S.intrinsics())   # close cameras: most of the cube is unseen
    marked = (m.density_grid < 0).cpu().numpy()
    assert 0 < marked.sum() < marked.size
    for it in range(3):
        m.iter_density = 16 if it else 0
        m.update_extra_state()
        g = m.density_grid.cpu().numpy()
        assert ((g < 0) == marked).all()
        ws = m._ngp_dg_ws
        mean, thresh = ws.state.cpu().numpy()
        assert abs(mean - m.mean_density) < 1e-12 + 1e-7 * abs(mean)
        assert thresh == min(mean, np.float32(m.density_thresh))
        assert (m.density_bitfield.cpu().numpy() == O.packbits(g, thresh)).all()
    pose = S.make_cameras(1, seed=3)[0]
    ro, rd = S.get_rays(pose, S.intrinsics(), 800, 800, torch.randint(0, 640000, (4096,), generator=gen(1)), device="cuda")
    nears, fars = raymarching.near_far_from_aabb(ro, rd, m.aabb_train, m.min_near)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(ro, rd, 1, m.density_bitfield, 1, 128, nears, fars, None, -1, False, 128,
                                                            True, 0, 1024)
    assert xyzs.shape[0] > 0
