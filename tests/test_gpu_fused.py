"""GPU: the fused field path (nerf_fused.fused_field) against the module-by-module path of network_ff.py."""
import numpy as np
import pytest
import torch

from util import gen, rel_err, synth_rays

pytestmark = pytest.mark.gpu


def _models():
    from nerf_step import NeRFFieldFF
    import ngp_synth as S
    torch.manual_seed(1)
    a = NeRFFieldFF(bound=1, fused=False).cuda().train()
    with torch.no_grad():
        a.encoder.embeddings.uniform_(-0.5, 0.5)
    grid, _ = S.box_union_density(128, seed=12)
    a.density_bitfield.copy_(torch.from_numpy(S.packbits_np(grid.numpy())).cuda())
    b = NeRFFieldFF(bound=1, fused=True).cuda().train()
    b.load_state_dict(a.state_dict())
    return a, b


@pytest.mark.parametrize("M", [1000, 128 * 300 + 17])
def test_fused_forward_matches_unfused(M):
    a, b = _models()
    x = (torch.rand(M, 3, generator=gen(1)) * 2.2 - 1.1).cuda()      # a few points outside the box -> zero features
    d = torch.randn(M, 3, generator=gen(2)); d = (d / d.norm(dim=-1, keepdim=True)).cuda()
    with torch.autocast("cuda", dtype=torch.float16):
        s0, c0 = a(x, d)
        s1, c1 = b(x, d)
    assert s1.dtype == torch.float32 and c1.dtype == torch.float32 and c1.shape == (M, 3)
    # same kernels, same rounding points: sigma within an ulp of exp, rgb within one fp16 ulp of sigmoid
    assert rel_err(s1.detach().cpu().numpy(), s0.detach().float().cpu().numpy()) < 1e-6
    assert (c1.detach() - c0.detach().float()).abs().max().item() <= 1e-3
    assert (c1.detach() != c0.detach().float()).float().mean().item() < 0.01


def test_fused_backward_matches_unfused():
    a, b = _models()
    M = 128 * 200 + 5
    x = (torch.rand(M, 3, generator=gen(3)) * 2 - 1).cuda()
    d = torch.randn(M, 3, generator=gen(4)); d = (d / d.norm(dim=-1, keepdim=True)).cuda()
    gs = torch.randn(M, generator=gen(5)).cuda() * 0.1
    gc = torch.randn(M, 3, generator=gen(6)).cuda()
    outs = []
    for m in (a, b):
        m.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            s, c = m(x, d)
        ((s.float() * gs).sum() + (c.float() * gc).sum()).backward()
        outs.append([p.grad.clone() for p in (m.encoder.embeddings, m.sigma_net.weights, m.color_net.weights)])

    def l2(u, v):
        return float((u - v).norm() / v.norm())
    # fp16 atomics on both sides (order-dependent) and half-rounded intermediate grads: Frobenius-norm agreement
    assert l2(outs[1][0], outs[0][0]) < 2e-2
    assert l2(outs[1][1], outs[0][1]) < 1e-2
    assert l2(outs[1][2], outs[0][2]) < 1e-2


def test_fused_train_step_runs_and_matches_loss():
    from nerf_step import train_step
    a, b = _models()
    N = 4096
    rays_o, rays_d, _, _ = synth_rays(N)
    ro, rd = rays_o.cuda(), rays_d.cuda()
    target = torch.rand(N, 3, generator=gen(5)).cuda()
    # loss scaling as in real training (GradScaler): without it the table gradients sit in fp16's subnormal range
    sa = torch.amp.GradScaler("cuda", init_scale=65536.0); sb = torch.amp.GradScaler("cuda", init_scale=65536.0)
    la, _ = train_step(a, ro, rd, target, None, sa, perturb=False, force_all_rays=True)
    lb, _ = train_step(b, ro, rd, target, None, sb, perturb=False, force_all_rays=True)
    assert abs(la.item() - lb.item()) < 1e-4 * max(1.0, abs(la.item()))
    ga, gb = a.encoder.embeddings.grad, b.encoder.embeddings.grad
    assert float((ga - gb).norm() / ga.norm()) < 3e-2


@pytest.mark.parametrize("M", [100, 128 * 2, 128 * 777 + 33])
def test_color_backward_dual_equals_single(M):
    """ngp_field_color_backward through the two-context kernel == through the single-context kernel (both output-gradient forms)."""
    import _ngp_b200 as nb
    lib = nb.load()
    a, _ = _models()
    d = torch.randn(M, 3, generator=gen(2)); d = (d / d.norm(dim=-1, keepdim=True)).cuda()
    h = (torch.randn(M, 16, generator=gen(3)) * 0.5).half().cuda()
    wc = a.color_net.weights.detach().half()
    nl = a.color_net.num_layers
    fb = torch.empty(nl, M, 64, dtype=torch.half, device="cuda"); rgb = torch.empty(M, 3, device="cuda")
    nb.call("ngp_field_color_forward", d.data_ptr(), h.data_ptr(), wc.data_ptr(), nl, M, 1, fb.data_ptr(), rgb.data_ptr())
    d_rgb = torch.randn(M, 3, generator=gen(4)).cuda(); d_sig = torch.randn(M, generator=gen(5)).cuda() * 0.1
    g3 = (torch.randn(M, 3, generator=gen(6)) * 0.1).half().cuda()
    nbytes = lib.ngp_ffmlp_backward_workspace_bytes(M, 32, 16, 64, nl)
    res = {}
    try:
        for dual in (0, 1):
            lib.ngp_debug_set_mlp_backward(dual)
            for form in ("rgb", "grad_h"):
                dys = torch.zeros(M, 16, dtype=torch.half, device="cuda"); gw = torch.zeros_like(wc); ws = torch.empty(nbytes // 4, device="cuda")
                if form == "rgb":
                    nb.call("ngp_field_color_backward", d_rgb.data_ptr(), rgb.data_ptr(), d_sig.data_ptr(), h.data_ptr(), d.data_ptr(), wc.data_ptr(),
                            fb.data_ptr(), nl, M, dys.data_ptr(), gw.data_ptr(), ws.data_ptr(), nbytes)
                else:
                    nb.call("ngp_field_color_backward_ex", None, None, g3.data_ptr(), None, h.data_ptr(), d.data_ptr(), None, wc.data_ptr(),
                            fb.data_ptr(), nl, M, dys.data_ptr(), gw.data_ptr(), ws.data_ptr(), nbytes, 0)
                torch.cuda.synchronize()
                res[(dual, form)] = (dys.clone(), ws.clone())
    finally:
        lib.ngp_debug_set_mlp_backward(1)
    for form in ("rgb", "grad_h"):
        assert torch.equal(res[(0, form)][0], res[(1, form)][0])
        assert rel_err(res[(1, form)][1].cpu().numpy(), res[(0, form)][1].cpu().numpy()) < 1e-5
        assert float(res[(1, form)][0].float().abs().sum()) > 0


def test_sigma_forward_gather_variants_are_bit_identical():
    """The two measured experiments of the fused encoder -> sigma kernel (aligned x-pair 8-byte gathers; coarse table levels staged in shared
    memory by cp.async.bulk) deliver the same table entries as the default 4-byte gathers: features, hidden stash, h and sigma identical."""
    import _ngp_b200 as nb
    from nerf_fused import field_forward, field_cfg
    a, b = _models()
    M = 128 * 37 + 5
    x = (torch.rand(M, 3, generator=gen(7)) * 2.1 - 1.05).cuda()
    d = torch.randn(M, 3, generator=gen(8)); d = (d / d.norm(dim=-1, keepdim=True)).cuda()
    cfg = field_cfg(b.encoder, b.sigma_net, b.color_net, b.bound, True)
    lib = nb.load()
    outs = []
    try:
        for variant, levels in ((0, 1), (1, 1), (2, 1), (2, 2)):
            assert lib.ngp_debug_set_sigma_gather(variant, levels) == 0
            sigma, rgb, stash = field_forward(x, d, b.encoder.embeddings, b.encoder.offsets, b.sigma_net.weights, b.color_net.weights, cfg)
            torch.cuda.synchronize()
            t = stash["tensors"]
            outs.append((sigma.clone(), rgb.clone(), t[5].clone(), t[6].clone(), t[8].clone()))     # sigma, rgb, feat, h, sigma-net stash
    finally:
        lib.ngp_debug_set_sigma_gather(-1, 1)
    for o in outs[1:]:
        for ref, got in zip(outs[0], o):
            assert torch.equal(ref, got)
    assert float(outs[0][2].abs().max()) > 0
