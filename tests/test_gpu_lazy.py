"""GPU: deferred tensors (ngp_lazy) — the fused encoder->MLP / SH-concat->MLP kernels reached through the UNCHANGED module sequence of
nerf/network_ff.py:51-74 give the same values and gradients as the literal op-by-op sequence, and every other consumer of a deferred
tensor sees exactly the eager result."""
import numpy as np
import pytest
import torch

from util import gen, rel_err

pytestmark = pytest.mark.gpu


def _field(seed=1):
    from nerf_step import NeRFFieldFF
    torch.manual_seed(seed)
    m = NeRFFieldFF(bound=1, fused=False).cuda().train()
    with torch.no_grad():
        m.encoder.embeddings.uniform_(-0.3, 0.3)
    return m


def _inputs(M, seed=0):
    x = (torch.rand(M, 3, generator=gen(seed)) * 2 - 1).cuda()
    d = torch.nn.functional.normalize(torch.randn(M, 3, generator=gen(seed + 1)), dim=-1).cuda()
    return x, d


def _run(m, x, d, lazy):
    import ngp_lazy
    import _ngp_b200 as nb
    ngp_lazy.enabled = lazy
    try:
        m.zero_grad(set_to_none=True)
        nb.profile_begin()
        with torch.autocast("cuda", dtype=torch.float16):
            sigma, rgb = m(x, d)
            loss = ((sigma.float() * 0.01).sum() + (rgb.float() * torch.linspace(0.5, 1.5, 3, device=x.device)).sum()) / x.shape[0]
        loss.backward()
        calls = [r[0] for r in nb.profile_end()]
        g = [p.grad.detach().float().clone() for p in (m.encoder.embeddings, m.sigma_net.weights, m.color_net.weights)]
        return sigma.detach().float(), rgb.detach().float(), g, calls
    finally:
        ngp_lazy.enabled = True


def test_network_ff_sequence_fuses_and_matches_eager():
    m = _field()
    x, d = _inputs(20000)          # ragged: not a multiple of 128
    s0, c0, g0, calls0 = _run(m, x, d, lazy=False)
    s1, c1, g1, calls1 = _run(m, x, d, lazy=True)
    # the literal sequence launches the separate encoder / SH / MLP ops; the deferred one only the fused kernels
    assert "ngp_grid_encode_forward" in calls0 and "ngp_sh_encode_forward" in calls0 and "ngp_ffmlp_forward" in calls0
    assert "ngp_grid_encode_forward" not in calls1 and "ngp_sh_encode_forward" not in calls1 and "ngp_ffmlp_forward" not in calls1
    assert "ngp_field_sigma_forward" in calls1 and "ngp_field_color_forward_ex" in calls1 and "ngp_field_color_backward_ex" in calls1
    # same encoder arithmetic (bit-identical features), same MLP kernels: sigma identical, rgb within fp16 rounding of the SH staging
    assert torch.equal(s0, s1)
    assert float((c0 - c1).abs().max()) <= 2e-3 and float(((c0 - c1).abs() > 0).float().mean()) < 0.05
    for a, b, tol in zip(g0, g1, (3e-2, 1e-2, 1e-2)):          # table: fp16 atomics order
        assert torch.isfinite(a).all() and torch.isfinite(b).all() and float(a.abs().max()) > 0
        assert rel_err(b.cpu().numpy(), a.cpu().numpy()) < tol


def test_inference_and_density_paths():
    import ngp_lazy
    m = _field().eval()
    x, d = _inputs(4099, seed=3)
    outs = []
    for lazy in (False, True):
        ngp_lazy.enabled = lazy
        try:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                sigma, rgb = m(x, d)
                den = m.density(x)
            outs.append((sigma.float(), rgb.float(), den["sigma"].float(), den["geo_feat"].float()))
        finally:
            ngp_lazy.enabled = True
    for a, b in zip(*outs):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= 2e-3 * max(1.0, float(a.abs().max()))
    assert torch.equal(outs[0][2], outs[1][2])


def test_other_consumers_get_the_eager_tensor():
    """nn.Linear / torch functions / indexing on a deferred encoder output (nerf/network.py:33-47 style) and a plain SH use."""
    import ngp_lazy
    from gridencoder import GridEncoder
    from shencoder import SHEncoder
    enc = GridEncoder(desired_resolution=2048).cuda()
    sh = SHEncoder().cuda()
    lin = torch.nn.Linear(32, 8, bias=False).cuda()
    x, d = _inputs(1000, seed=5)
    with torch.autocast("cuda", dtype=torch.float16):
        f = enc(x, bound=1)
        assert isinstance(f, ngp_lazy.DeferredGridFeatures) and tuple(f.shape) == (1000, 32) and f.dtype == torch.half and f.is_cuda
        y = lin(f)                                   # materialises
        e = enc._forward_eager(x, 1)
        assert torch.equal(y, lin(e)) and torch.equal(f[3:5], e[3:5]) and torch.equal(f.float().sum(0), e.float().sum(0))
        y.float().sum().backward()
        assert enc.embeddings.grad is not None and float(enc.embeddings.grad.abs().sum()) > 0
        s = sh(d)
        assert isinstance(s, ngp_lazy.DeferredSH)
        assert torch.equal(s * 1.0, sh._forward_eager(d, 1))
        # a cat that is not the network_ff pattern falls back to the plain op
        c = torch.cat([sh(d), torch.ones(1000, 4, device="cuda")], dim=-1)
        assert c.shape == (1000, 20) and not isinstance(c, ngp_lazy.Deferred)
    # outside autocast (fp32 table) and with coordinate gradients the encoder stays eager
    assert not isinstance(enc(x, bound=1), ngp_lazy.Deferred)
    xg = x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        assert not isinstance(enc(xg, bound=1), ngp_lazy.Deferred)


def test_sdf_backbone_fuses():
    """sdf/netowrk_ff.py:37-46: GridEncoder -> FFMLP(32 -> 64 x3 -> 1)."""
    import ngp_lazy
    import _ngp_b200 as nb
    from gridencoder import GridEncoder
    from ffmlp import FFMLP
    torch.manual_seed(0)
    enc = GridEncoder(desired_resolution=2048).cuda()
    with torch.no_grad():
        enc.embeddings.uniform_(-0.3, 0.3)
    net = FFMLP(32, 1, 64, 3).cuda().train()
    x, _ = _inputs(30000, seed=9)
    res = []
    for lazy in (False, True):
        ngp_lazy.enabled = lazy
        try:
            enc.zero_grad(); net.zero_grad()
            nb.profile_begin()
            with torch.autocast("cuda", dtype=torch.float16):
                h = net(enc(x))
                assert tuple(h.shape) == (30000, 1)
                (h.float().abs().sum() * 16).backward()
            calls = [r[0] for r in nb.profile_end()]
            res.append((h.detach().float(), enc.embeddings.grad.clone() / 16, net.weights.grad.clone() / 16, calls))
        finally:
            ngp_lazy.enabled = True
    assert "ngp_field_sigma_forward" in res[1][3] and "ngp_field_sigma_forward" not in res[0][3]
    assert torch.equal(res[0][0], res[1][0])
    assert rel_err(res[1][1].cpu().numpy(), res[0][1].cpu().numpy()) < 3e-2
    assert rel_err(res[1][2].cpu().numpy(), res[0][2].cpu().numpy()) < 1e-2
