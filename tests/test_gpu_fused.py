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
