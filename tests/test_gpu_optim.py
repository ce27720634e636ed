"""GPU: fused optimizer step (csrc/optim.cu) against torch.optim.Adam + torch.amp.GradScaler semantics."""
import numpy as np
import pytest
import torch

from util import gen, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [100003, 100000])      # scalar tail path / 128-bit vector path
def test_adam_matches_torch_and_skips_on_inf(n):
    import _ngp_b200 as nb
    p0 = torch.randn(n, generator=gen(1)).cuda()
    p = p0.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    shadow = torch.empty(n, dtype=torch.half, device="cuda")
    state = torch.zeros(8, dtype=torch.int32, device="cuda"); state[0:1].view(torch.float32).fill_(1024.0)
    state[4:5].view(torch.float32).fill_(1.0)          # lr_scale
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    for it in range(5):
        g = torch.randn(n, generator=gen(10 + it)).cuda() * 0.01
        inject_inf = (it == 2)
        g16 = (g * 1024.0 * (0.5 if it > 2 else 1.0)).half()       # after the skipped step the scale is halved
        if inject_inf:
            g16[123] = float("inf")
        else:
            ref.grad = g16.float() / (1024.0 * (0.5 if it > 2 else 1.0))
            opt.step()
        gbuf = g16.clone()
        nb.call("ngp_optim_check_finite", gbuf.data_ptr(), 1, n, state.data_ptr())
        nb.call("ngp_optim_adam_step", p.data_ptr(), m.data_ptr(), v.data_ptr(), gbuf.data_ptr(), 1, shadow.data_ptr(), n,
                1e-2, 0.9, 0.99, 1e-15, state.data_ptr(), 1)
        nb.call("ngp_optim_scaler_update", state.data_ptr(), 2.0, 0.5, 2000)
        assert float(gbuf.abs().max()) == 0.0                                  # gradient consumed and zeroed
        st = state.cpu()
        assert st[2].item() == 0
        assert st[3].item() == (it + 1 if it < 2 else it)                       # the inf step was skipped
        assert st[0:1].view(torch.float32).item() == (1024.0 if it < 2 else 512.0)
        assert rel_err(p.cpu().numpy(), ref.detach().cpu().numpy()) < 1e-5
    assert torch.equal(shadow, p.half())


def test_fused_optimizer_trains_the_field():
    """End to end: fp16 gradient sink + fused Adam reduce the loss like GradScaler + torch Adam do."""
    from nerf_step import NeRFFieldFF
    from ngp_optim import FusedFieldOptimizer
    import ngp_synth as S
    from util import synth_rays
    torch.manual_seed(1)
    model = NeRFFieldFF(bound=1, fused=True).cuda().train()
    grid, _ = S.box_union_density(128, seed=12)
    model.density_bitfield.copy_(torch.from_numpy(S.packbits_np(grid.numpy())).cuda())
    opt = FusedFieldOptimizer(model.encoder, model.sigma_net, model.color_net, lr=1e-2, init_scale=1024.0)
    N = 4096
    rays_o, rays_d, _, _ = synth_rays(N)
    ro, rd = rays_o.cuda(), rays_d.cuda()
    target = torch.rand(1, 3).expand(N, 3).contiguous().cuda()
    losses = []
    for it in range(12):
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.render_train(ro, rd, perturb=True, force_all_rays=True)
            loss = torch.nn.functional.mse_loss(out["image"], target)
        (loss * opt.scale_tensor()).backward()
        assert model.encoder.embeddings.grad is None            # no fp32 .grad was materialised
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    assert int(opt.state[3].item()) == 12
    # the fp16 table shadow the next forward will read is the one the kernel refreshed
    from gridencoder.grid import _half_table
    assert torch.equal(_half_table(model.encoder.embeddings), model.encoder.embeddings.detach().half())


def test_fused_train_step_matches_autograd_path():
    """FusedTrainStep (closed-form MSE gradient, no autograd) == autograd through the fused field + same optimizer."""
    from nerf_step import NeRFFieldFF, FusedTrainStep
    from ngp_optim import FusedFieldOptimizer
    import ngp_synth as S
    from util import synth_rays
    N = 8192
    rays_o, rays_d, _, _ = synth_rays(N)
    ro, rd = rays_o.cuda(), rays_d.cuda()
    target = torch.rand(N, 3, generator=gen(5)).cuda()
    grid, _ = S.box_union_density(128, seed=12)
    res = []
    for manual in (False, True):
        torch.manual_seed(1)
        model = NeRFFieldFF(bound=1, fused=True).cuda().train()
        with torch.no_grad():
            model.encoder.embeddings.uniform_(-0.3, 0.3)
        model.density_bitfield.copy_(torch.from_numpy(S.packbits_np(grid.numpy())).cuda())
        opt = FusedFieldOptimizer(model.encoder, model.sigma_net, model.color_net, lr=1e-2, init_scale=1024.0)
        model.mean_count = 200000
        if manual:
            loss = FusedTrainStep(model, opt, N, perturb=False)(ro, rd, target)
        else:
            with torch.autocast("cuda", dtype=torch.float16):
                out = model.render_train(ro, rd, perturb=False)
                loss = ((out["image"] - target) ** 2).sum() / (3.0 * N)
            (loss * opt.scale_tensor()).backward()
            opt.step()
        res.append((loss.item(), model.encoder.embeddings.detach().clone(), model.color_net.weights.detach().clone(), int(opt.state[3].item())))
    (la, ea, wa, sa), (lb, eb, wb, sb) = res
    assert sa == 1 and sb == 1
    assert abs(la - lb) < 1e-6 * max(1.0, abs(la))
    # Adam's first step moves every touched parameter by ~lr*sign(g): compare the update directions
    ua, ub = ea - ea.mean() * 0, eb
    assert float((ea - eb).abs().max()) <= 2.5e-2                      # at most a sign flip on near-zero gradients (2*lr)
    assert float(((ea - eb).abs() > 1e-3).float().mean()) < 5e-3       # ...and only for a tiny fraction of entries
    assert float((wa - wb).abs().max()) <= 2.5e-2


def test_resume_from_reference_adam_state():
    """FusedFieldOptimizer.load_state_dict(torch Adam state in the reference's group layout) then one fused step == torch Adam's
    next step from the same state (checkpoint interchange, nerf/utils.py:1118-1137)."""
    from nerf_step import NeRFFieldFF
    from ngp_optim import FusedFieldOptimizer
    torch.manual_seed(0)
    a = NeRFFieldFF().cuda()
    b = NeRFFieldFF().cuda()
    b.load_state_dict(a.state_dict())
    pa = [a.encoder.embeddings, a.sigma_net.weights, a.color_net.weights]
    ref = torch.optim.Adam([{"params": [pa[0]]}, {"params": [pa[1]]}, {"params": []}, {"params": [pa[2]]}], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    grads = [[(torch.randn(p.shape, generator=gen(50 + 3 * it + i)) * 1e-3).cuda() for i, p in enumerate(pa)] for it in range(4)]
    for it in range(3):
        for p, g in zip(pa, grads[it]):
            p.grad = g.clone()
        ref.step()
    b.load_state_dict(a.state_dict())
    fused = FusedFieldOptimizer(b.encoder, b.sigma_net, b.color_net)
    fused.load_state_dict(ref.state_dict())
    fused.load_scaler_state_dict({"scale": 1024.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 0})
    for p, g in zip(pa, grads[3]):
        p.grad = (g * 1024.0).half().float() / 1024.0          # what the fp16 sink can represent
    ref.step()
    for (p, off, k), g in zip(fused.segments, grads[3]):
        fused.sink[off:off + k].copy_((g * 1024.0).half().reshape(-1))
    fused.step()
    pb = [b.encoder.embeddings, b.sigma_net.weights, b.color_net.weights]
    for x, y in zip(pa, pb):
        assert rel_err(y.detach().cpu().numpy(), x.detach().cpu().numpy()) < 1e-5
    sd = fused.state_dict()
    assert float(sd["state"][0]["step"]) == 4.0
    assert rel_err(sd["state"][2]["exp_avg_sq"].cpu().numpy(), ref.state_dict()["state"][2]["exp_avg_sq"].cpu().numpy()) < 1e-5
    fused.detach()


def _fresh_field(grid):
    from nerf_step import NeRFFieldFF
    from ngp_optim import FusedFieldOptimizer
    import ngp_synth as S
    torch.manual_seed(1)
    model = NeRFFieldFF(bound=1, fused=True).cuda().train()
    with torch.no_grad():
        model.encoder.embeddings.uniform_(-0.3, 0.3)
    model.density_bitfield.copy_(torch.from_numpy(S.packbits_np(grid.numpy())).cuda())
    opt = FusedFieldOptimizer(model.encoder, model.sigma_net, model.color_net, lr=1e-2, init_scale=1024.0)
    model.mean_count = 200000
    return model, opt


def test_pipelined_prefetched_steps_match_sequential_and_capture():
    """step_prefetched (march of step i+1 on the prefetch stream, 4 field chunks on the side stream) == the plain sequential step;
    the same call captured in two CUDA graphs (even / odd sample slot) and replayed continues the sequence."""
    from nerf_step import FusedTrainStep
    import ngp_synth as S
    from util import synth_rays
    N, K = 8192, 4
    grid, _ = S.box_union_density(128, seed=12)
    rays = [tuple(t.cuda() for t in synth_rays(N, seed=s)[:2]) for s in range(K + 1)]
    tgts = [torch.rand(N, 3, generator=gen(40 + s)).cuda() for s in range(K + 1)]
    # A: sequential, unchunked
    ma, oa = _fresh_field(grid)
    fa = FusedTrainStep(ma, oa, N, perturb=False, chunks=1)
    la = [float(fa(rays[i][0], rays[i][1], tgts[i])) for i in range(K)]
    # B: prefetched + chunked, eager
    mb, ob = _fresh_field(grid)
    fb = FusedTrainStep(mb, ob, N, perturb=False, chunks=4)
    fb.march(*rays[0])
    lb = [float(fb.step_prefetched(tgts[i], *rays[i + 1])) for i in range(K)]
    assert int(ob.state[3].item()) == K and int(oa.state[3].item()) == K
    # same samples (bit-exact marcher), same arithmetic up to fp16 atomics order; Adam's early sign-like steps amplify that a little
    for x, y in zip(la, lb):
        assert abs(x - y) <= 2e-3 * max(abs(x), 1e-6), (la, lb)
    assert la[-1] < la[0]
    assert fb.sync_host_state() == K + 1 and mb.local_step == K + 1          # K + 1 marches were recorded in the ring
    # C: the same call under CUDA-graph capture (what bench.py replays)
    mc, oc = _fresh_field(grid)
    fc = FusedTrainStep(mc, oc, N, perturb=False, chunks=4)
    st_o, st_d, st_t = torch.empty_like(rays[0][0]), torch.empty_like(rays[0][1]), torch.empty_like(tgts[0])
    hp = torch.cuda.Stream(priority=-1)
    hp.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(hp):
        fc.march(*rays[0])
        lc = [float(fc.step_prefetched(tgts[i], *rays[i + 1])) for i in range(2)]        # eager warm-up: steps 0, 1
    torch.cuda.current_stream().wait_stream(hp)
    torch.cuda.synchronize()
    graphs, losses = [], []
    for par in range(2):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=hp):
            losses.append(fc.step_prefetched(st_t, st_o, st_d))
        graphs.append(g)
    torch.cuda.synchronize()
    for i in range(2, K):
        st_t.copy_(tgts[i]); st_o.copy_(rays[i + 1][0]); st_d.copy_(rays[i + 1][1])
        graphs[i % 2].replay()
        lc.append(float(losses[i % 2]))
    assert int(oc.state[3].item()) == K
    for x, y in zip(la, lc):
        assert abs(x - y) <= 2e-3 * max(abs(x), 1e-6), (la, lc)
