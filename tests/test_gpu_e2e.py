"""GPU: the whole hot path through the drop-in packages (march -> encode -> MLP -> composite -> backward)."""
import numpy as np
import pytest
import torch

from util import gen, rel_err, synth_rays

pytestmark = pytest.mark.gpu


def _model():
    from nerf_step import NeRFFieldFF
    import ngp_synth as S
    torch.manual_seed(1)
    m = NeRFFieldFF(bound=1).cuda()
    grid, _ = S.box_union_density(128, seed=12)
    m.density_bitfield.copy_(torch.from_numpy(S.packbits_np(grid.numpy())).cuda())
    return m


def test_state_dict_contract():
    m = _model()
    sd = m.state_dict()
    assert tuple(sd["encoder.embeddings"].shape) == (6119864, 2) and sd["encoder.embeddings"].dtype == torch.float32
    assert tuple(sd["encoder.offsets"].shape) == (17,) and sd["encoder.offsets"].dtype == torch.int32
    assert tuple(sd["sigma_net.weights"].shape) == (7168,) and tuple(sd["color_net.weights"].shape) == (11264,)


def test_train_step_vs_oracle_pipeline():
    """One training forward+backward: image vs an fp32 torch re-evaluation of the same samples (the marcher
    output is taken from the GPU, its parity is covered by test_gpu_raymarching)."""
    from nerf_step import train_step
    from oracle import oracle as O
    m = _model().train()
    with torch.no_grad():
        m.encoder.embeddings.uniform_(-0.5, 0.5)
    N = 4096
    rays_o, rays_d, _, _ = synth_rays(N)
    ro, rd = rays_o.cuda(), rays_d.cuda()
    target = torch.rand(N, 3, generator=gen(5)).cuda()
    loss, out = train_step(m, ro, rd, target, perturb=False, force_all_rays=True)
    assert torch.isfinite(loss) and out["n_samples"] > N
    assert m.encoder.embeddings.grad is not None and m.encoder.embeddings.grad.abs().sum() > 0
    assert m.sigma_net.weights.grad.abs().sum() > 0 and m.color_net.weights.grad.abs().sum() > 0
    # re-evaluate the field in fp32 torch on the marched samples and composite with the oracle
    import raymarching
    with torch.no_grad():
        nears, fars = raymarching.near_far_from_aabb(ro, rd, m.aabb_train, m.min_near)
        xyzs, dirs, deltas, rays = raymarching.march_rays_train(ro, rd, 1, m.density_bitfield, 1, 128, nears, fars, None, -1, False, 128, True, 0, 1024)
        with torch.autocast("cuda", dtype=torch.float16):
            sig, rgb = m(xyzs, dirs)
        x01 = ((xyzs + 1) / 2).cpu().numpy()
        S_ = float(np.log2(m.encoder.per_level_scale))
        feat = O.grid_forward(x01, m.encoder.embeddings.half().cpu().numpy(), m.encoder.offsets.cpu().numpy(), S_, 16)
        pad = (-len(feat)) % 128
        y, _ = O.mlp_forward(np.concatenate([feat, np.zeros((pad, 32), np.float16)]),
                             m.sigma_net.weights.half().cpu().numpy(), 32, 64, 2)
        y = y[:len(feat)].astype(np.float32)
        sig_ref = np.exp(y[:, 0])
        assert rel_err(sig.float().cpu().numpy(), sig_ref) < 5e-3
        ws, dp, im = O.composite_rays_train_forward(sig.float().cpu().numpy(), rgb.float().cpu().numpy(), deltas.cpu().numpy(), rays.cpu().numpy(), 1e-4)
        img = im + (1 - ws)[:, None]
    assert rel_err(out["image"].detach().float().cpu().numpy(), img) < 1e-4


def test_steady_state_mean_count_path_and_optimizer():
    """mean_count fast path (no D2H sync inside the step) + Adam + GradScaler run and reduce the loss."""
    from nerf_step import train_step
    m = _model().train()
    N = 4096
    rays_o, rays_d, _, _ = synth_rays(N)
    ro, rd = rays_o.cuda(), rays_d.cuda()
    target = torch.rand(1, 3).expand(N, 3).contiguous().cuda()
    opt = torch.optim.Adam(m.parameters(), lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    losses = []
    for it in range(12):
        if it % 4 == 0:
            m.update_mean_count()
        loss, out = train_step(m, ro, rd, target, opt, scaler, perturb=True)
        losses.append(loss.item())
    assert m.mean_count > 0
    assert losses[-1] < losses[0]
