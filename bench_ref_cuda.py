"""bench_ref_cuda.py — time the REFERENCE's own CUDA extensions (oracle/_ref, unmodified sources built for sm_100) on
the same training step as bench.py, on the same GPU.  Reported beside our number as `ref_cuda` (the north_star's
">= 10x the reference's own CUDA-extension build" denominator).  Measurement infrastructure only.

The reference's Python wrappers cannot travel to the GPU box, so the autograd glue below issues the same native calls
with the same tensor preparation as gridencoder/grid.py:24-90, ffmlp/ffmlp.py:15-83,147-168,
shencoder/sphere_harmonics.py:14-54 and raymarching/raymarching.py:161-291 (including their permute / cat-pad /
zeros_like copies), and the same caller sequence as nerf/network_ff.py:51-74 + nerf/renderer.py:280-321.
"""
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from oracle import ref_driver as R


class _Grid(Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, pls, H):
        inputs = inputs.contiguous()
        emb = embeddings.to(torch.half)          # grid.py:43-44: cast on every call
        out, _ = R.grid_encode_forward(inputs, emb, offsets, pls, H)
        ctx.save_for_backward(inputs, emb, offsets)
        ctx.cfg = (pls, H)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, emb, offsets = ctx.saved_tensors
        ge, _ = R.grid_encode_backward(grad.contiguous(), inputs, emb, offsets, ctx.cfg[0], ctx.cfg[1])
        return None, ge, None, None, None


class _MLP(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.half)
    def forward(ctx, inputs, weights, ind, nl, calc_gi):
        inputs = inputs.contiguous(); weights = weights.contiguous()
        out, fb = R.ffmlp_forward(inputs, weights, ind, 16, 64, nl)
        ctx.save_for_backward(inputs, weights, fb)
        ctx.cfg = (ind, nl, calc_gi)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, weights, fb = ctx.saved_tensors
        ind, nl, calc_gi = ctx.cfg
        gi, gw, _ = R.ffmlp_backward(grad.contiguous(), inputs, weights, fb, ind, 16, 64, nl, calc_grad_inputs=calc_gi)
        return gi, gw, None, None, None


def _mlp(x, w, ind, nl, out_dim):
    B = x.shape[0]
    pad = 128 - (B % 128)                       # ffmlp.py:157-159
    x = torch.cat([x, torch.zeros(pad, x.shape[1], dtype=x.dtype, device=x.device)], dim=0)
    return _MLP.apply(x, w, ind, nl, x.requires_grad)[:B, :out_dim]


class _Composite(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T):
        sigmas = sigmas.contiguous(); rgbs = rgbs.contiguous()
        ws, dp, im = R.composite_rays_train_forward(sigmas, rgbs, deltas, rays, T)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, im)
        ctx.T = T
        return ws, dp, im

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, gws, gdp, gim):
        sigmas, rgbs, deltas, rays, ws, im = ctx.saved_tensors
        gs, gc = R.composite_rays_train_backward(gws.contiguous(), gim.contiguous(), sigmas, rgbs, deltas, rays, ws, im, ctx.T)
        return gs, gc, None, None, None


def measure(dev, R_rays, dev_in, model, steps=5, warmup=2):
    """Same parameters / bitfield / rays as `model` (our NeRFFieldFF); returns rays/s of the reference CUDA build."""
    if not R.available():
        return {"unavailable": "oracle/_ref not built"}
    from nerf_step import trunc_exp
    emb = model.encoder.embeddings.detach().clone().requires_grad_(True)
    ws_ = model.sigma_net.weights.detach().clone().requires_grad_(True)
    wc_ = model.color_net.weights.detach().clone().requires_grad_(True)
    offsets = model.encoder.offsets
    pls = model.encoder.per_level_scale
    opt = torch.optim.Adam([emb, ws_, wc_], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda", init_scale=128.0)
    M = int(model.mean_count); M += 128 - M % 128
    N = dev_in[0][0].shape[0]
    counter = torch.zeros(2, dtype=torch.int32, device=dev)

    def step(ro, rd, tgt):
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            nears, fars = R.near_far_from_aabb(ro, rd, model.aabb_train, model.min_near)
            counter.zero_()
            noises = torch.rand(N, device=dev)
            xyzs, dirs, deltas, rays, _ = R.march_rays_train(ro, rd, 1.0, model.density_bitfield, 1, 128, nears, fars, M, noises, 0.0, 1024, counter)
            x = (xyzs + 1) / 2
            h = _mlp(_Grid.apply(x, emb, offsets, pls, 16), ws_, 32, 2, 16)
            sigma = trunc_exp(h[..., 0]); geo = h[..., 1:]
            d, _ = R.sh_encode_forward(dirs, 4)
            hc = torch.cat([d, geo, torch.zeros_like(geo[..., :1])], dim=-1)
            rgb = torch.sigmoid(_mlp(hc, wc_, 32, 3, 3))
            wsum, depth, image = _Composite.apply(sigma, rgb, deltas, rays, 1e-4)
            image = image + (1 - wsum).unsqueeze(-1)
            loss = ((image - tgt) ** 2).sum() / (3.0 * R_rays)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        return loss

    for i in range(warmup):
        step(*dev_in[i % len(dev_in)])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(*dev_in[i % len(dev_in)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": R_rays / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms, "steps": steps, "loss": float(loss),
            "what": "unmodified reference CUDA extensions (gridencoder, ffmlp+CUTLASS 2.8, shencoder, raymarching) built for sm_100, same step, same GPU"}


if __name__ == "__main__":
    import argparse, json, os, sys
    ROOT = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200")]
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays-per-step", type=int, default=640000)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import bench
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    model, _ = bench.build_model(dev)
    _, dev_in = bench.make_inputs(a.rays_per_step, 0, 1, dev)
    # steady-state sample budget, measured with the reference marcher itself
    nears, fars = R.near_far_from_aabb(dev_in[0][0], dev_in[0][1], model.aabb_train, model.min_near)
    counts = []
    for ro, rd, _t in dev_in:
        n_, f_ = R.near_far_from_aabb(ro, rd, model.aabb_train, model.min_near)
        c = torch.zeros(2, dtype=torch.int32, device=dev)
        mod = R.mod("raymarching")
        N = ro.shape[0]
        e = torch.empty(0, 3, device=dev); e2 = torch.empty(0, 2, device=dev)
        mod.march_rays_train(ro, rd, model.density_bitfield, 1.0, 0.0, 1024, N, 1, 128, 0, n_, f_, e, e, e2,
                             torch.empty(N, 3, dtype=torch.int32, device=dev), c, torch.ones(N, device=dev))
        counts.append(int(c[0].item()))
    model.mean_count = max(counts)
    print(json.dumps(measure(dev, a.rays_per_step, dev_in, model, steps=a.steps)))
