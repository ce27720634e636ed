"""GPU parity: occupancy-grid marcher + compositor (csrc/raymarch.cu) vs the C oracle and oracle/_ref."""
import numpy as np
import pytest
import torch

from util import gen, rel_err, canon_rays, gather_segments, synth_rays

pytestmark = pytest.mark.gpu


def _scene(N, seed=3, perturb=True):
    rays_o, rays_d, bitfield, grid = synth_rays(N, seed)
    noises = torch.rand(N, generator=gen(4)) if perturb else torch.zeros(N)
    return rays_o, rays_d, bitfield, grid, noises


def test_near_far_morton_packbits_bitexact():
    from oracle import oracle as O
    import raymarching
    rays_o, rays_d, bitfield, grid, _ = _scene(5000)
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0])
    # include rays that miss the box
    rays_d[:100] = -rays_d[:100]
    n, f = raymarching.near_far_from_aabb(rays_o.cuda(), rays_d.cuda(), aabb.cuda(), 0.2)
    rn, rf = O.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb.numpy(), 0.2)
    np.testing.assert_array_equal(n.cpu().numpy(), rn)
    np.testing.assert_array_equal(f.cpu().numpy(), rf)
    coords = torch.randint(0, 128, (4096, 3), generator=gen(1), dtype=torch.int32)
    idx = raymarching.morton3D(coords.cuda())
    np.testing.assert_array_equal(idx.cpu().numpy(), O.morton3D(coords.numpy()))
    back = raymarching.morton3D_invert(idx)
    np.testing.assert_array_equal(back.cpu().numpy(), coords.numpy())
    g = torch.rand(2, 128 ** 3, generator=gen(2))
    bits = raymarching.packbits(g.cuda(), 0.5)
    np.testing.assert_array_equal(bits.cpu().numpy(), O.packbits(g.numpy(), 0.5))


@pytest.mark.parametrize("dt_gamma", [0.0, 1.0 / 128])
@pytest.mark.parametrize("perturb", [False, True])
def test_march_rays_train_bitexact_vs_oracle(dt_gamma, perturb):
    """Per-ray sample counts, positions and deltas are bit-exact against the CPU restatement."""
    from oracle import oracle as O
    import _ngp_b200 as nb
    N = 4096
    rays_o, rays_d, bitfield, grid, noises = _scene(N, perturb=perturb)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = O.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb, 0.2)
    M = N * 256
    ox, od_, odl, orays, ocnt = O.march_rays_train(rays_o.numpy(), rays_d.numpy(), bitfield.numpy(), 1.0, dt_gamma, 1024,
                                                   1, 128, M, nears, fars, noises.numpy())
    dev = "cuda"
    xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev); deltas = torch.zeros(M, 2, device=dev)
    rays = torch.zeros(N, 3, dtype=torch.int32, device=dev); counter = torch.zeros(2, dtype=torch.int32, device=dev)
    args = [t.cuda() for t in (rays_o, rays_d, bitfield)]
    nd, fd, nz = torch.from_numpy(nears).cuda(), torch.from_numpy(fars).cuda(), noises.cuda()
    nb.call("ngp_march_rays_train", args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), 1.0, dt_gamma, 1024, N, 1,
            128, M, nd.data_ptr(), fd.data_ptr(), xyzs.data_ptr(), dirs.data_ptr(), deltas.data_ptr(), rays.data_ptr(),
            counter.data_ptr(), nz.data_ptr())
    rays_h = rays.cpu().numpy()
    assert counter.cpu().tolist() == ocnt.tolist()
    assert ocnt[0] > N  # the scene is not empty
    if dt_gamma == 0.0:
        assert (orays[:, 2] > 64).any() and (orays[:, 2] <= 64).any()   # both the smem-cached and the fallback emission paths run
    np.testing.assert_array_equal(canon_rays(rays_h)[:, [0, 2]], canon_rays(orays)[:, [0, 2]])   # ids + counts
    # offsets form a permutation-free tiling of [0, total)
    r = rays_h[np.argsort(rays_h[:, 1], kind="stable")]
    r = r[r[:, 2] > 0]
    assert r[0, 1] == 0 and np.all(r[1:, 1] == r[:-1, 1] + r[:-1, 2])
    for a, b in ((xyzs, ox), (dirs, od_), (deltas, odl)):
        np.testing.assert_array_equal(gather_segments(a.cpu().numpy(), rays_h), gather_segments(b, orays))


def test_march_multi_cascade_bitexact_vs_oracle():
    """bound = 2 -> two cascades (mip level selection by position / step size, raymarching.cu:42-54, 368)."""
    from oracle import oracle as O
    import _ngp_b200 as nb
    N = 2048
    rays_o, rays_d, _, _, noises = _scene(N)
    rays_o = rays_o * 1.6
    bits = (torch.rand(2 * 128 ** 3 // 8, generator=gen(9)) < 0.02).to(torch.uint8) * torch.randint(1, 255, (2 * 128 ** 3 // 8,), generator=gen(10), dtype=torch.uint8)
    aabb = np.array([-2, -2, -2, 2, 2, 2], np.float32)
    nears, fars = O.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb, 0.2)
    for dt_gamma in (0.0, 1.0 / 128):
        M = N * 512
        ox, od_, odl, orays, ocnt = O.march_rays_train(rays_o.numpy(), rays_d.numpy(), bits.numpy(), 2.0, dt_gamma, 1024, 2, 128, M,
                                                       nears, fars, noises.numpy())
        xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
        rays = torch.zeros(N, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
        ro, rd, bf = rays_o.cuda(), rays_d.cuda(), bits.cuda()
        nd, fd, nz = torch.from_numpy(nears).cuda(), torch.from_numpy(fars).cuda(), noises.cuda()
        nb.call("ngp_march_rays_train", ro.data_ptr(), rd.data_ptr(), bf.data_ptr(), 2.0, dt_gamma, 1024, N, 2, 128, M,
                nd.data_ptr(), fd.data_ptr(), xyzs.data_ptr(), dirs.data_ptr(), deltas.data_ptr(), rays.data_ptr(),
                counter.data_ptr(), nz.data_ptr())
        assert counter.cpu().tolist() == ocnt.tolist() and ocnt[0] > 0
        r = rays.cpu().numpy()
        np.testing.assert_array_equal(canon_rays(r)[:, [0, 2]], canon_rays(orays)[:, [0, 2]])
        np.testing.assert_array_equal(gather_segments(xyzs.cpu().numpy(), r), gather_segments(ox, orays))
        np.testing.assert_array_equal(gather_segments(deltas.cpu().numpy(), r), gather_segments(odl, orays))


def test_march_small_max_steps_clamp_order():
    """max_steps < H / 2^(C-1) makes dt_min > dt_max; the reference's clamp fminf(max, fmaxf(min, x)) then returns dt_max.  With
    dt_gamma == 0 the marcher's constant step must be that value in the probe loop AND in the emitted deltas (raymarching.cu:345-398)."""
    from oracle import oracle as O
    import _ngp_b200 as nb
    N, max_steps = 2048, 64
    rays_o, rays_d, bitfield, grid, noises = _scene(N)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = O.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb, 0.2)
    M = N * 128
    ox, od_, odl, orays, ocnt = O.march_rays_train(rays_o.numpy(), rays_d.numpy(), bitfield.numpy(), 1.0, 0.0, max_steps, 1, 128, M,
                                                   nears, fars, noises.numpy())
    xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
    rays = torch.zeros(N, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    ro, rd, bf = rays_o.cuda(), rays_d.cuda(), bitfield.cuda()
    nd, fd, nz = torch.from_numpy(nears).cuda(), torch.from_numpy(fars).cuda(), noises.cuda()
    nb.call("ngp_march_rays_train", ro.data_ptr(), rd.data_ptr(), bf.data_ptr(), 1.0, 0.0, max_steps, N, 1, 128, M,
            nd.data_ptr(), fd.data_ptr(), xyzs.data_ptr(), dirs.data_ptr(), deltas.data_ptr(), rays.data_ptr(),
            counter.data_ptr(), nz.data_ptr())
    assert counter.cpu().tolist() == ocnt.tolist() and ocnt[0] > 0
    r = rays.cpu().numpy()
    np.testing.assert_array_equal(canon_rays(r)[:, [0, 2]], canon_rays(orays)[:, [0, 2]])
    np.testing.assert_array_equal(gather_segments(xyzs.cpu().numpy(), r), gather_segments(ox, orays))
    np.testing.assert_array_equal(gather_segments(deltas.cpu().numpy(), r), gather_segments(odl, orays))
    dt_max = np.float32(2 * 1.7320508075688772) * np.float32(1.0) / np.float32(128)
    assert np.all(gather_segments(deltas.cpu().numpy(), r)[:, 0] == dt_max)


def test_march_rays_train_overflow_and_wrapper():
    """M smaller than the total: dropped rays keep (id, offset, count) but write nothing; wrapper shape rules."""
    import raymarching
    N = 2048
    rays_o, rays_d, bitfield, grid, _ = _scene(N)
    ro, rd, bf = rays_o.cuda(), rays_d.cuda(), bitfield.cuda()
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0]).cuda()
    nears, fars = raymarching.near_far_from_aabb(ro, rd, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(ro, rd, 1, bf, 1, 128, nears, fars, counter, -1, False, 128, False, 0, 1024)
    total = int(counter[0].item())
    assert int(counter[1].item()) == N and xyzs.shape[0] % 128 == 0 and xyzs.shape[0] >= total and xyzs.shape[0] - total <= 128
    assert rays[:, 2].sum().item() == total
    # fast path with an under-estimated budget
    counter.zero_()
    m = max(128, total // 2)
    x2, d2, l2, r2 = raymarching.march_rays_train(ro, rd, 1, bf, 1, 128, nears, fars, counter, m, False, 128, False, 0, 1024)
    M = x2.shape[0]
    assert M == m + (128 - m % 128)
    assert int(counter[0].item()) == total
    dropped = (r2[:, 1] + r2[:, 2] > M) & (r2[:, 2] > 0)
    assert dropped.any()
    kept = r2[~dropped & (r2[:, 2] > 0)]
    assert (l2[kept[0, 1].item(), 0] > 0)
    # empty input
    e = raymarching.march_rays_train(ro[:0], rd[:0], 1, bf, 1, 128, nears[:0], fars[:0], None, -1, False, 128, True, 0, 1024)
    assert e[3].shape == (0, 3)


def test_composite_train_vs_oracle_and_gradcheck():
    from oracle import oracle as O
    import raymarching
    N = 1024
    rays_o, rays_d, bitfield, grid, noises = _scene(N)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = O.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb, 0.2)
    M = N * 256
    ox, od_, odl, orays, ocnt = O.march_rays_train(rays_o.numpy(), rays_d.numpy(), bitfield.numpy(), 1.0, 0.0, 1024, 1, 128,
                                                   M, nears, fars, noises.numpy())
    m = int(ocnt[0])
    sig = (torch.rand(m, generator=gen(20)) * 30).requires_grad_(True)
    rgb = torch.rand(m, 3, generator=gen(21)).requires_grad_(True)
    dl = torch.from_numpy(odl[:m].copy()); rays_t = torch.from_numpy(orays)
    s_d, c_d = sig.detach().cuda().requires_grad_(True), rgb.detach().cuda().requires_grad_(True)
    ws, depth, image = raymarching.composite_rays_train(s_d, c_d, dl.cuda(), rays_t.cuda(), 1e-4)
    ows, odepth, oimage = O.composite_rays_train_forward(sig.detach().numpy(), rgb.detach().numpy(), dl.numpy(), orays, 1e-4)
    # __expf vs expf: 1e-5 of the output scale
    assert rel_err(ws.detach().cpu().numpy(), ows) < 1e-5
    assert rel_err(depth.detach().cpu().numpy(), odepth) < 1e-5
    assert rel_err(image.detach().cpu().numpy(), oimage) < 1e-5
    gws = torch.randn(N, generator=gen(22)); gim = torch.randn(N, 3, generator=gen(23))
    (ws * gws.cuda()).sum().backward(retain_graph=True)
    gs1 = s_d.grad.clone(); s_d.grad = None; c_d.grad = None
    ((ws * gws.cuda()).sum() + (image * gim.cuda()).sum()).backward()
    ogs, ogc = O.composite_rays_train_backward(gws.numpy(), gim.numpy(), sig.detach().numpy(), rgb.detach().numpy(),
                                               dl.numpy(), orays, ows, oimage, 1e-4)
    assert rel_err(s_d.grad.cpu().numpy(), ogs) < 1e-4
    assert rel_err(c_d.grad.cpu().numpy(), ogc) < 1e-5
    assert gs1.abs().sum().item() > 0


def test_composite_forward_mse_head_and_counter_push():
    """Step-driver extensions: the compositor with the fused loss head (ngp_composite_rays_train_forward_mse) against the plain compositor
    followed by the torch expressions it replaces, and the device-side sample-count ring (ngp_step_counter_push)."""
    from oracle import oracle as O
    import raymarching
    import _ngp_b200 as nb
    N = 2048
    rays_o, rays_d, bitfield, grid, noises = _scene(N)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = O.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb, 0.2)
    M = N * 64                                  # some rays overflow the budget and composite to the background
    ox, od_, odl, orays, ocnt = O.march_rays_train(rays_o.numpy(), rays_d.numpy(), bitfield.numpy(), 1.0, 0.0, 1024, 1, 128,
                                                   M, nears, fars, noises.numpy())
    m = min(int(ocnt[0]), M)
    sig = torch.zeros(M); sig[:m] = torch.rand(m, generator=gen(20)) * 30
    rgb = torch.zeros(M, 3); rgb[:m] = torch.rand(m, 3, generator=gen(21))
    dl = torch.zeros(M, 2); dl[:m] = torch.from_numpy(odl[:m].copy())
    sig, rgb, dl, rays_t = sig.cuda(), rgb.cuda(), dl.cuda(), torch.from_numpy(orays).cuda()
    target = torch.rand(N, 3, generator=gen(24)).cuda()
    ws, depth, image = raymarching.composite_rays_train(sig, rgb, dl, rays_t, 1e-4)
    bg, R, scale = 1.0, float(4 * N), torch.tensor([128.0], device="cuda")
    pred = image + (1 - ws).unsqueeze(-1) * bg
    diff = pred - target
    g_pred = diff * ((2.0 / (3.0 * R)) * scale)
    g_ws = -(g_pred.sum(-1)) * bg
    loss = (diff * diff).sum() / (3.0 * R)
    f = dict(dtype=torch.float32, device="cuda")
    ws2, dp2, im2 = torch.empty(N, **f), torch.empty(N, **f), torch.empty(N, 3, **f)
    gi, gw, sq = torch.empty(N, 3, **f), torch.empty(N, **f), torch.empty(N, **f)
    nb.call("ngp_composite_rays_train_forward_mse", sig.data_ptr(), rgb.data_ptr(), dl.data_ptr(), rays_t.data_ptr(), M, N, 1e-4,
            target.data_ptr(), bg, float(2.0 / (3.0 * R)), scale.data_ptr(), ws2.data_ptr(), dp2.data_ptr(), im2.data_ptr(), gi.data_ptr(),
            gw.data_ptr(), sq.data_ptr())
    assert torch.equal(ws2, ws) and torch.equal(dp2, depth) and torch.equal(im2, image)      # same compositor arithmetic
    assert rel_err(gi.cpu().numpy(), g_pred.cpu().numpy()) < 1e-6
    assert rel_err(gw.cpu().numpy(), g_ws.cpu().numpy()) < 1e-6
    assert abs(float(sq.sum() / (3.0 * R)) - float(loss)) < 1e-6 * float(loss)
    # the ring: three pushes land in rows 0, 1, 2; row 15 wraps to 0
    ring = torch.zeros(1, dtype=torch.int32, device="cuda"); nsteps = torch.zeros(1, dtype=torch.int32, device="cuda")
    sc = torch.zeros(16, 2, dtype=torch.int32, device="cuda")
    for k in range(18):
        counter = torch.tensor([100 + k, 7 + k], dtype=torch.int32, device="cuda")
        nb.call("ngp_step_counter_push", ring.data_ptr(), counter.data_ptr(), nsteps.data_ptr(), sc.data_ptr())
    torch.cuda.synchronize()
    assert int(nsteps) == 18 and int(ring) == 2
    assert sc[0].tolist() == [116, 23] and sc[1].tolist() == [117, 24] and sc[2].tolist() == [102, 9] and sc[15].tolist() == [115, 22]


def test_inference_march_and_composite_vs_oracle():
    from oracle import oracle as O
    import raymarching
    N = 3000
    rays_o, rays_d, bitfield, grid, _ = _scene(N)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = O.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb, 0.2)
    alive = np.arange(N, dtype=np.int32)[::2].copy()
    n_alive, n_step = len(alive), 4
    noises = np.zeros(n_alive, np.float32)
    ox, od_, odl = O.march_rays(n_alive, n_step, alive, nears, rays_o.numpy(), rays_d.numpy(), 1.0, 0.0, 1024, 1, 128,
                                bitfield.numpy(), nears, fars, noises, align=128)
    nd, fd = torch.from_numpy(nears).cuda(), torch.from_numpy(fars).cuda()
    rt = nd.clone(); al = torch.from_numpy(alive).cuda()
    x, d, l = raymarching.march_rays(n_alive, n_step, al, rt, rays_o.cuda(), rays_d.cuda(), 1, bitfield.cuda(), 1, 128, nd, fd, 128, False, 0, 1024)
    for a, b in ((x, ox), (d, od_), (l, odl)):
        np.testing.assert_array_equal(a.cpu().numpy(), b)
    M = x.shape[0]
    sig = torch.rand(M, generator=gen(30)) * 40; rgb = torch.rand(M, 3, generator=gen(31))
    ws = torch.zeros(N); dp = torch.zeros(N); im = torch.zeros(N, 3)
    o_alive, o_t, o_ws, o_dp, o_im = O.composite_rays(n_alive, n_step, 1e-2, alive, nears, sig.numpy(), rgb.numpy(), odl,
                                                      ws.numpy(), dp.numpy(), im.numpy())
    wsd, dpd, imd = ws.cuda(), dp.cuda(), im.cuda()
    raymarching.composite_rays(n_alive, n_step, al, rt, sig.cuda(), rgb.cuda(), l, wsd, dpd, imd, 1e-2)
    np.testing.assert_array_equal(al.cpu().numpy(), o_alive)   # which rays terminated: integer parity
    assert rel_err(rt.cpu().numpy(), o_t) < 1e-6
    assert rel_err(wsd.cpu().numpy(), o_ws) < 1e-5 and rel_err(imd.cpu().numpy(), o_im) < 1e-5 and rel_err(dpd.cpu().numpy(), o_dp) < 1e-5


def test_vs_reference_extension():
    """Identical inputs through the reference's own raymarching kernels: counts / positions bit-exact,
    compositor outputs and gradients equal up to fp32 re-association (same __expf, warp-scan summation order)."""
    from oracle import ref_driver as R
    if not R.available("raymarching"):
        pytest.skip("oracle/_ref/raymarching not built")
    import _ngp_b200 as nb
    import raymarching
    N = 100000
    rays_o, rays_d, bitfield, grid, noises = _scene(N)
    ro, rd, bf, nz = rays_o.cuda(), rays_d.cuda(), bitfield.cuda(), noises.cuda()
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0]).cuda()
    rn, rf = R.near_far_from_aabb(ro, rd, aabb, 0.2)
    n, f = raymarching.near_far_from_aabb(ro, rd, aabb, 0.2)
    assert torch.equal(n, rn) and torch.equal(f, rf)
    for dt_gamma in (0.0, 1 / 256):
        M = N * 128
        rx, rdd, rl, rr, rc = R.march_rays_train(ro, rd, 1.0, bf, 1, 128, rn, rf, M, nz, dt_gamma, 1024)
        xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
        rays = torch.zeros(N, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
        nb.call("ngp_march_rays_train", ro.data_ptr(), rd.data_ptr(), bf.data_ptr(), 1.0, dt_gamma, 1024, N, 1, 128, M,
                rn.data_ptr(), rf.data_ptr(), xyzs.data_ptr(), dirs.data_ptr(), deltas.data_ptr(), rays.data_ptr(),
                counter.data_ptr(), nz.data_ptr())
        assert torch.equal(counter, rc)
        a, b = rays.cpu().numpy(), rr.cpu().numpy()
        np.testing.assert_array_equal(canon_rays(a)[:, [0, 2]], canon_rays(b)[:, [0, 2]])
        np.testing.assert_array_equal(gather_segments(xyzs.cpu().numpy(), a), gather_segments(rx.cpu().numpy(), b))
        np.testing.assert_array_equal(gather_segments(deltas.cpu().numpy(), a), gather_segments(rl.cpu().numpy(), b))
    m = int(counter[0].item())
    sig = (torch.rand(M, generator=gen(40)) * 30).cuda(); rgb = torch.rand(M, 3, generator=gen(41)).cuda()
    rws, rdp, rim = R.composite_rays_train_forward(sig, rgb, deltas, rays, 1e-4)
    ws, dp, im = raymarching.composite_rays_train(sig, rgb, deltas, rays, 1e-4)
    # the compositor evaluates the per-ray recurrence with warp prefix scans: same terms, re-associated fp32 sums
    assert rel_err(ws.cpu().numpy(), rws.cpu().numpy()) < 1e-5 and rel_err(im.cpu().numpy(), rim.cpu().numpy()) < 1e-5
    assert rel_err(dp.cpu().numpy(), rdp.cpu().numpy()) < 1e-5
    gws = torch.randn(N, device="cuda"); gim = torch.randn(N, 3, device="cuda")
    rgs, rgc = R.composite_rays_train_backward(gws, gim, sig, rgb, deltas, rays, rws, rim, 1e-4)
    gs = torch.zeros_like(sig); gc = torch.zeros_like(rgb)
    nb.call("ngp_composite_rays_train_backward", gws.data_ptr(), gim.data_ptr(), sig.data_ptr(), rgb.data_ptr(),
            deltas.data_ptr(), rays.data_ptr(), rws.data_ptr(), rim.data_ptr(), M, N, 1e-4, gs.data_ptr(), gc.data_ptr())
    assert rel_err(gs.cpu().numpy(), rgs.cpu().numpy()) < 1e-4 and rel_err(gc.cpu().numpy(), rgc.cpu().numpy()) < 1e-5
    # inference pair
    alive = torch.arange(N, dtype=torch.int32, device="cuda"); rt = rn.clone(); nz0 = torch.zeros(N, device="cuda")
    rx, rdd, rl = R.march_rays(N, 2, alive, rt, ro, rd, 1.0, bf, 1, 128, rn, rf, nz0, 128)
    x, d, l = raymarching.march_rays(N, 2, alive, rt, ro, rd, 1, bf, 1, 128, rn, rf, 128, False, 0, 1024)
    assert torch.equal(x, rx) and torch.equal(l, rl) and torch.equal(d, rdd)


def test_time_indexed_bitfield_slices():
    """D-NeRF keeps [T, ...] occupancy state and hands slices to the same ops (dnerf/renderer.py:92-93, 295, 362, 546-547):
    packbits must write through a slice view, the marchers must read one."""
    import raymarching
    import ngp_synth as S
    T, H = 4, 128
    grid, _ = S.box_union_density(H, seed=12)
    grids = torch.zeros(T, 1, H ** 3, device="cuda")
    grids[2] = grid.cuda()
    bits = torch.zeros(T, H ** 3 // 8, dtype=torch.uint8, device="cuda")
    for t in range(T):
        out = raymarching.packbits(grids[t], 0.01, bits[t])
        assert out.data_ptr() == bits[t].data_ptr()
    ref = torch.from_numpy(S.packbits_np(grid.numpy())).cuda()
    assert torch.equal(bits[2], ref) and int(bits[0].sum()) == 0 and int(bits[3].sum()) == 0
    rays_o, rays_d, _, _ = synth_rays(2048)
    ro, rd = rays_o.cuda(), rays_d.cuda()
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0], device="cuda")
    nears, fars = raymarching.near_far_from_aabb(ro, rd, aabb, 0.2)
    a = raymarching.march_rays_train(ro, rd, 1, bits[2], 1, H, nears, fars, None, -1, False, 128, True, 0, 1024)
    b = raymarching.march_rays_train(ro, rd, 1, ref.clone(), 1, H, nears, fars, None, -1, False, 128, True, 0, 1024)
    assert a[0].shape == b[0].shape and a[0].shape[0] > 0
    assert (gather_segments(a[0].cpu().numpy(), a[3].cpu().numpy()) == gather_segments(b[0].cpu().numpy(), b[3].cpu().numpy())).all()
    e = raymarching.march_rays_train(ro, rd, 1, bits[0], 1, H, nears, fars, None, -1, False, 128, True, 0, 1024)
    assert int(e[3][:, 2].sum()) == 0                     # empty time slice: no samples
