"""GPU: the remaining BASELINE.json configs as parity cases (config 3: fused density inference on 4096 rays x 1024
samples; config 5: SDF field, hashgrid + FFMLP(32-64-64-64-1), 1M surface points)."""
import numpy as np
import pytest
import torch

from util import gen, rel_err

pytestmark = pytest.mark.gpu


def test_config3_fused_density_inference_4096x1024():
    """4 194 304 points along 4096 rays clipped to the AABB; model.eval(); fused encoder->sigma-MLP inference kernel vs the
    module-by-module path (GridEncoder -> FFMLP inference -> trunc_exp) at full size, and vs the CPU oracle on a slice."""
    from nerf_step import NeRFFieldFF
    from oracle import oracle as O
    import _ngp_b200 as nb
    from gridencoder.grid import _half_table
    from util import synth_rays
    torch.manual_seed(1)
    m = NeRFFieldFF(bound=1).cuda().eval()
    with torch.no_grad():
        m.encoder.embeddings.uniform_(-0.5, 0.5)
    rays_o, rays_d, _, _ = synth_rays(4096, seed=6)
    t = torch.linspace(2.0, 4.5, 1024)
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * t[None, :, None]).clamp(-1, 1).reshape(-1, 3).cuda()
    M = pts.shape[0]
    assert M == 4096 * 1024
    import ngp_lazy
    ngp_lazy.enabled = False
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            ref = m.density(pts)["sigma"]                              # literal module sequence, inference kernels
    finally:
        ngp_lazy.enabled = True
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        via_modules = m.density(pts)["sigma"]                          # the same call, deferred-tensor fusion on (default)
    assert torch.equal(via_modules, ref)
    x01 = ((pts + 1) / 2).contiguous()
    table = _half_table(m.encoder.embeddings)
    h = torch.empty(M, 16, dtype=torch.half, device="cuda"); sig = torch.empty(M, device="cuda")
    w = m.sigma_net.weights.detach().half()
    nb.call("ngp_field_sigma_forward", x01.data_ptr(), table.data_ptr(), m.encoder.offsets.data_ptr(),
            16, float(np.log2(m.encoder.per_level_scale)), 16, 0, 0, w.data_ptr(), 2, M, 0, None, None, h.data_ptr(), sig.data_ptr())
    assert rel_err(sig.cpu().numpy(), ref.float().cpu().numpy()) < 1e-6
    # oracle on the first 2048 points
    n = 2048
    feat = O.grid_forward(x01[:n].cpu().numpy(), m.encoder.embeddings.detach().half().cpu().numpy(), m.encoder.offsets.cpu().numpy(),
                          float(np.log2(m.encoder.per_level_scale)), 16)
    y, _ = O.mlp_forward(feat, w.cpu().numpy(), 32, 64, 2)
    assert rel_err(h[:n].float().cpu().numpy(), y.astype(np.float32)) < 3e-3


def test_config5_sdf_field_1m_points():
    """sdf/netowrk_ff.py:9-46 topology: GridEncoder(16 levels, ->2048) -> FFMLP(32 -> 64 x3 -> 1), 2^20 points near a sphere,
    mape-style loss; forward vs oracle on a slice, gradients finite and adjoint-consistent at full size."""
    from gridencoder import GridEncoder
    from ffmlp import FFMLP
    from oracle import oracle as O
    torch.manual_seed(2)
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048).cuda()
    with torch.no_grad():
        enc.embeddings.uniform_(-0.1, 0.1)
    backbone = FFMLP(input_dim=32, output_dim=1, hidden_dim=64, num_layers=3).cuda().train()
    B = 1 << 20
    g = gen(7)
    d = torch.randn(B, 3, generator=g); d = d / d.norm(dim=-1, keepdim=True)
    X = (d * 0.8 + torch.randn(B, 3, generator=g) * 0.01)
    X[B * 7 // 8:] = torch.rand(B - B * 7 // 8, 3, generator=g) * 2 - 1
    X = X.clamp(-1, 1).cuda()
    target = (X.norm(dim=-1, keepdim=True) - 0.8)
    with torch.autocast("cuda", dtype=torch.float16):
        y = backbone(enc(X, bound=1))
        assert y.shape == (B, 1) and y.dtype == torch.float16
        loss = (torch.abs(y.float() - target) / (torch.abs(target) + 1e-2)).mean()
    (loss * 1024).backward()
    assert torch.isfinite(loss)
    ge, gw = enc.embeddings.grad, backbone.weights.grad
    assert torch.isfinite(ge).all() and torch.isfinite(gw).all() and ge.abs().sum() > 0 and gw.abs().sum() > 0
    # forward vs oracle on a slice
    n = 4096
    S = float(np.log2(enc.per_level_scale))
    feat = O.grid_forward(((X[:n] + 1) / 2).cpu().numpy(), enc.embeddings.detach().half().cpu().numpy(), enc.offsets.cpu().numpy(), S, 16)
    yo, _ = O.mlp_forward(feat, backbone.weights.detach().half().cpu().numpy(), 32, 64, 3)
    assert rel_err(y[:n, 0].float().detach().cpu().numpy(), yo[:, 0].astype(np.float32)) < 5e-3
    # size-independent property: only table entries reachable from the sampled points received gradient (<= 128 per point)
    assert int((ge.abs().sum(-1) > 0).sum()) <= B * 128
