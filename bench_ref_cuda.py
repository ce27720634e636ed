"""bench_ref_cuda.py — time the REFERENCE on the same training step as bench.py, on the same GPU: the reference's own Python
(nerf/network_ff.py NeRFNetwork + nerf/renderer.py run_cuda, its wrapper packages gridencoder/ffmlp/shencoder/raymarching — unmodified
copies staged by oracle/build_ref.ship_python) over the reference's own CUDA extensions (oracle/_ref/*.so, built from
/root/reference/*/src with the single flag change -std=c++17).  Reported beside our number as `ref_cuda` (the north_star's
">= 10x the reference's own CUDA-extension build (--fp16 --ff --cuda_ray)" denominator).  Measurement infrastructure only.

One step = the reference trainer's iteration (nerf/utils.py:861-868): optimizer.zero_grad -> autocast{ model.render(...) -> MSE } ->
scaler.scale(loss).backward() -> scaler.step(Adam) -> scaler.update(), with the steady-state sample budget (`mean_count`) the
reference would have reached after its first epoch.  Before timing, one un-perturbed step from identical parameters is run through
BOTH stacks (reference callers over this repo's drop-in packages, reference callers over the reference's extensions) and the two
losses are compared (`loss_match`).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200")]


def _adopt(model, src):
    """copy parameters / occupancy of our bench model (nerf_step.NeRFFieldFF) into a reference NeRFNetwork"""
    with torch.no_grad():
        model.encoder.embeddings.copy_(src.encoder.embeddings)
        model.sigma_net.weights.copy_(src.sigma_net.weights)
        model.color_net.weights.copy_(src.color_net.weights)
        model.density_bitfield.copy_(src.density_bitfield)
    model.mean_count = int(src.mean_count)


def first_step_loss(stack, src, ro, rd, tgt, R_rays):
    """loss of one un-perturbed training forward through the reference callers in `stack`"""
    from oracle import ref_stack
    model = ref_stack.make_nerf(stack, bound=1).cuda()
    _adopt(model, src)
    with stack.active():
        model.train()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            out = model.render(ro[None], rd[None], staged=False, bg_color=None, perturb=False, force_all_rays=False, dt_gamma=0, max_steps=1024)
            loss = ((out["image"][0] - tgt) ** 2).sum() / (3.0 * R_rays)
    return float(loss)


def measure(dev, R_rays, dev_in, src, steps=5, warmup=2, which="ref", lazy=True):
    """which = "ref": reference callers over the reference's wrappers + extensions (the `ref_cuda` arm);
    which = "ours": the SAME unmodified callers over this repo's drop-in packages (the `dropin` arm: what a torch-ngp user gets by
    swapping the four package directories and nothing else)."""
    from oracle import ref_stack
    if not (ref_stack.available("ref") and ref_stack.available("ours")):
        return {"unavailable": "oracle/_ref (reference extensions + staged reference Python) not built"}
    ref = ref_stack.load("ref")
    ours = ref_stack.load("ours")
    import ngp_lazy
    ngp_lazy.enabled = bool(lazy)
    # parity anchor: same parameters, same rays, no perturbation -> the two stacks must produce the same loss
    l_ref = first_step_loss(ref, src, *dev_in[0], R_rays)
    l_ours = first_step_loss(ours, src, *dev_in[0], R_rays)
    match = abs(l_ref - l_ours) <= 2e-3 * max(abs(l_ref), 1e-6)

    ref = ref if which == "ref" else ours
    model = ref_stack.make_nerf(ref, bound=1).cuda()
    _adopt(model, src)
    with ref.active():
        model.train()
        opt = torch.optim.Adam(model.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)      # main_nerf.py:132
        scaler = torch.amp.GradScaler("cuda", init_scale=128.0)

        def step(ro, rd, tgt):
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.float16):
                out = model.render(ro[None], rd[None], staged=False, bg_color=None, perturb=True, force_all_rays=False, dt_gamma=0,
                                   max_steps=1024)
                loss = ((out["image"][0] - tgt) ** 2).sum() / (3.0 * R_rays)
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
            "first_step_loss_reference_stack": l_ref, "first_step_loss_dropin_stack": l_ours, "loss_match": bool(match),
            "what": ("reference wrappers, unmodified: nerf/network_ff.py + nerf/renderer.py + gridencoder/ffmlp/shencoder/raymarching "
                     "Python of the reference over its own CUDA extensions (gridencoder, ffmlp+CUTLASS 2.8, shencoder, raymarching) built "
                     "for sm_100; trainer iteration of nerf/utils.py:861-868 (GradScaler + torch Adam), same step, same GPU") if which == "ref" else
                    ("reference callers, unmodified (nerf/network_ff.py + nerf/renderer.py + encoding.py + activation.py), over THIS repo's "
                     "drop-in packages (gridencoder, ffmlp, shencoder, raymarching); deferred-tensor fusion "
                     + ("on" if lazy else "off (literal op-by-op sequence)") + "; trainer iteration of nerf/utils.py:861-868 (GradScaler + "
                     "torch Adam), same step, same GPU")}


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays-per-step", type=int, default=640000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--stack", default="ref", choices=["ref", "ours"])
    ap.add_argument("--no-lazy", action="store_true")
    a = ap.parse_args()
    import bench
    import raymarching
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    src, _ = bench.build_model(dev)
    _, dev_in = bench.make_inputs(a.rays_per_step, 0, 1, dev)
    # steady-state sample budget (per-ray counts are bit-identical between the two marchers: tests/test_gpu_raymarching.py)
    counts = []
    for ro, rd, _t in dev_in:
        nears, fars = raymarching.near_far_from_aabb(ro, rd, src.aabb_train, src.min_near)
        c = torch.zeros(2, dtype=torch.int32, device=dev)
        raymarching.march_rays_train(ro, rd, src.bound, src.density_bitfield, src.cascade, src.grid_size, nears, fars, c, -1, True, 128,
                                     False, 0, 1024)
        counts.append(int(c[0].item()))
    src.mean_count = max(counts)
    torch.cuda.empty_cache()
    print(json.dumps(measure(dev, a.rays_per_step, dev_in, src, steps=a.steps, which=a.stack, lazy=not a.no_lazy)))
