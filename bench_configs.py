"""bench_configs.py — throughput arms for the BASELINE.json configs other than the training step (bench.py --config c1|c3|c5|infer).

    c1     GridEncoder fwd + bwd on 65 536 random 3-D points, L=16 F=2 T=2^19 (testing/test_hashencoder.py)      -> points/s
    c3     fused encoder -> sigma-MLP inference, 4096 rays x 1024 samples (NeRFNetwork.density, model.eval())      -> rays/s
    c5     SDF mode: hashgrid + 1-output FFMLP on 2^20 synthetic surface points, fwd + bwd (sdf/netowrk_ff.py)       -> points/s
    infer  full 800x800 eval render through run_cuda's inference branch (nerf/renderer.py:323-372)                  -> rays/s

Each arm prints ONE JSON line shaped like bench.py's: value (device-timed, inputs resident), e2e (host buffers, copies inside the
timed region), roofline of the dominant kernel (per-call CUDA events), cpu_baseline (oracle port on a bounded sample, host cores)
and `ref_cuda` = the reference's own wrappers + CUDA extensions (oracle/_ref) on the same inputs on the same GPU.  L2: every timed
iteration is preceded by an untimed 512 MB fill (inputs of c1 fit in L2), see `config.l2`.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200")]


def _flush_buf(dev):
    return torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def _time_iters(fn, iters, warmup, flush, sync_each=False):
    """mean ms of fn() over `iters` runs, each bracketed by its own CUDA events after an untimed L2 flush."""
    for _ in range(warmup):
        flush.fill_(1)
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


def _profile_calls(fn, flush):
    """per-C-ABI-call device time of one fn() (after a flush): {name: (ms, args)} summed over calls of the same name"""
    import _ngp_b200 as nb
    flush.fill_(1)
    torch.cuda.synchronize()
    nb.profile_begin()
    fn()
    torch.cuda.synchronize()
    rec = nb.profile_end()
    out = {}
    for name, a, s0, s1 in rec:
        d = out.setdefault(name, [0.0, a, 0])
        d[0] += s0.elapsed_time(s1); d[2] += 1
    return out, len(rec)


def _roof(name, ms, nbytes, flops, pk):
    gb = nbytes / (ms * 1e-3) / 1e9
    tf = flops / (ms * 1e-3) / 1e12
    return {"kernel": name, "bound": "hbm", "achieved": gb, "peak": pk["hbm"], "unit": "GB/s", "frac": gb / pk["hbm"],
            "tensor_tflops": tf or None, "tensor_frac_of_sustained": (tf / pk["tf_sust"]) if tf else None, "traffic": None,
            "peak_source": pk["src"], "avg_launch_ms": ms, "algorithmic_bytes_per_launch": nbytes}


def _clocks(sampler, t0, t1):
    return sampler.stop(t0, t1)


def _cores():
    return min(os.cpu_count() or 1, 32)


def _finish(line, args):
    main_mod = sys.modules.get("__main__")
    emit = getattr(main_mod, "emit", None)          # bench.py run as a script owns the saved stdout descriptor
    if emit is None:
        import bench
        emit = bench.emit
    emit(line)


# ================================================================================================ C1
def run_c1(args, dev, pk, sampler):
    import bench
    import _ngp_b200 as nb
    from gridencoder import GridEncoder
    from oracle import oracle as O, ref_driver as R
    B = 65536
    g = torch.Generator().manual_seed(0)
    x_host = torch.rand(B, 3, generator=g).pin_memory()
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048).to(dev)
    with torch.no_grad():
        enc.embeddings.copy_((torch.rand(enc.embeddings.shape, generator=torch.Generator().manual_seed(1)) * 2 - 1))
    gy_host = torch.randn(B, 32, generator=torch.Generator().manual_seed(2)).half().pin_memory()
    x = x_host.to(dev); gy = gy_host.to(dev)
    xw = (x * 2 - 1).contiguous()
    flush = _flush_buf(dev)
    import ngp_lazy
    ngp_lazy.enabled = False          # this config measures the stand-alone encoder op

    def fwd_bwd(xin=xw, gin=gy):
        enc.embeddings.grad = None
        with torch.autocast("cuda", dtype=torch.float16):
            y = enc(xin, bound=1)
        y.backward(gin)
        return y

    t0 = sampler.mark()
    ms = _time_iters(fwd_bwd, args.steps, max(args.warmup, 3), flush)
    t1 = sampler.mark()
    calls, nlaunch = _profile_calls(fwd_bwd, flush)
    k_f, k_b = calls["ngp_grid_encode_forward"], calls["ngp_grid_encode_backward"]
    # e2e: host points + host upstream gradient in, the 24.5 MB fp16 table gradient stays on the device (it feeds the optimizer); a 4-byte
    # checksum of the features comes back
    xs = torch.empty_like(x); gs = torch.empty_like(gy)

    def e2e():
        xs.copy_(x_host, non_blocking=True); gs.copy_(gy_host, non_blocking=True)
        y = fwd_bwd((xs * 2 - 1), gs)
        return float(y.float().sum().item())
    ms_e2e = _time_iters(e2e, args.steps, 2, flush)
    # reference extension, same inputs (its wrapper: per-call table cast, [L,B,C] output + permute, zeros_like + permute in the backward)
    ref = None
    if R.available("gridencoder"):
        emb16 = None

        def ref_fb():
            e16 = enc.embeddings.detach().to(torch.half)         # grid.py:43-44
            out, _ = R.grid_encode_forward(x, e16, enc.offsets, enc.per_level_scale, 16)
            R.grid_encode_backward(gy, x, e16, enc.offsets, enc.per_level_scale, 16)
        ms_ref = _time_iters(ref_fb, max(3, min(args.steps, 10)), 2, flush)
        ref = {"value": B / (ms_ref * 1e-3), "unit": "points/s", "ms_per_iter": ms_ref,
               "what": "reference gridencoder extension (oracle/_ref) driven as gridencoder/grid.py:24-90 does, same inputs, same GPU"}
    # CPU: the C oracle port, single thread + OpenMP as built, bounded sample = the full 65 536 points once
    tab16 = enc.embeddings.detach().half().cpu().numpy()
    S = float(np.log2(enc.per_level_scale))
    tc = time.time()
    O.grid_forward(x_host.numpy(), tab16, enc.offsets.cpu().numpy(), S, 16)
    O.grid_backward(gy_host.numpy(), x_host.numpy(), enc.offsets.cpu().numpy(), tab16.shape[0], 2, S, 16)
    cpu_s = time.time() - tc
    nbytes = B * 588
    line = {"metric": "GridEncoder fwd+bwd points/sec (device-timed), 65536 random 3-D points, L=16 F=2 T=2^19 fp16 table", "value": B / (ms * 1e-3),
            "unit": "points/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "c1_gridencoder_fwd_bwd_64k_points", "points": B, "l2": "flushed (512 MB fill) before every timed iteration",
                       "note": "one step = GridEncoder forward + backward (table gradient) under autocast, as testing/test_hashencoder.py"},
            "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "points/s", "h2d_bytes_per_step": B * 12 + B * 64, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
            "gpu_launches": nlaunch, "clocks": _clocks(sampler, t0, t1),
            "roofline": _roof("ngp_grid_encode_backward", k_b[0], nbytes, 0, pk),
            "kernels": {"ngp_grid_encode_forward": {"ms": k_f[0], "GBps": nbytes / (k_f[0] * 1e-3) / 1e9},
                        "ngp_grid_encode_backward": {"ms": k_b[0], "GBps": nbytes / (k_b[0] * 1e-3) / 1e9}},
            "cpu_baseline": {"value": B / cpu_s, "unit": "points/s", "cores": int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)), "kind": "port",
                             "sample": f"oracle/ngp_oracle.c grid forward + backward on all 65536 points once ({cpu_s:.2f} s)"},
            "ref_cuda": ref}
    _finish(line, args)


# ================================================================================================ C3
def _density_points(dev):
    """4096 random rays of the synthetic camera x 1024 uniformly spaced samples, clipped to the AABB (seed 6)."""
    import ngp_synth as S
    g = torch.Generator().manual_seed(6)
    inds = torch.randint(0, 800 * 800, (4096,), generator=g)
    rays_o, rays_d = S.get_rays(S.make_cameras(4, seed=11)[6 % 4], S.intrinsics(), 800, 800, inds)
    t = torch.linspace(2.0, 4.5, 1024)
    return (rays_o[:, None, :] + rays_d[:, None, :] * t[None, :, None]).clamp(-1, 1).reshape(-1, 3)


def run_c3(args, dev, pk, sampler):
    import bench
    from oracle import oracle as O, ref_stack
    src, _ = bench.build_model(dev)
    with torch.no_grad():
        src.encoder.embeddings.uniform_(-0.5, 0.5)
    pts_host = _density_points(dev).contiguous().pin_memory()
    pts = pts_host.to(dev)
    M, NR = pts.shape[0], 4096
    flush = _flush_buf(dev)
    ours = ref_stack.load("ours") if ref_stack.available("ours") else None
    if ours is not None:
        model = ref_stack.make_nerf(ours, bound=1).to(dev).eval()       # the reference's own NeRFNetwork over our packages
        model.load_state_dict(src.state_dict(), strict=False)
        ctx = ours.active
        who = "reference nerf/network_ff.py NeRFNetwork.density (unmodified) over this repo's packages"
    else:
        import contextlib
        model, ctx, who = src.eval(), contextlib.nullcontext, "nerf_step.NeRFFieldFF.density (restated caller)"

    def dens(p=pts):
        with ctx(), torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return model.density(p)["sigma"]

    t0 = sampler.mark()
    ms = _time_iters(dens, args.steps, max(args.warmup, 3), flush)
    t1 = sampler.mark()
    calls, nlaunch = _profile_calls(dens, flush)
    k = calls.get("ngp_field_sigma_forward")
    stage = torch.empty_like(pts)

    def e2e():
        stage.copy_(pts_host, non_blocking=True)
        return float(dens(stage).sum().item())
    ms_e2e = _time_iters(e2e, max(3, args.steps // 2), 1, flush)
    ref = None
    if ref_stack.available("ref"):
        rs = ref_stack.load("ref")
        rm = ref_stack.make_nerf(rs, bound=1).to(dev).eval()
        rm.load_state_dict(src.state_dict(), strict=False)

        def rdens():
            with rs.active(), torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                return rm.density(pts)["sigma"]
        ms_ref = _time_iters(rdens, max(3, min(args.steps, 10)), 2, flush)
        err = float((rdens().float() - dens().float()).abs().max() / rdens().float().abs().max())
        ref = {"value": NR / (ms_ref * 1e-3), "unit": "rays/s", "ms_per_iter": ms_ref, "max_rel_diff_vs_ours": err,
               "what": "reference NeRFNetwork.density over the reference's wrappers + CUDA extensions (oracle/_ref), same points, same GPU"}
    # CPU oracle on a bounded sample (64 rays x 1024 samples)
    n = 64 * 1024
    tab16 = src.encoder.embeddings.detach().half().cpu().numpy()
    S = float(np.log2(src.encoder.per_level_scale))
    w = src.sigma_net.weights.detach().half().cpu().numpy()
    tc = time.time()
    feat = O.grid_forward(((pts_host[:n] + 1) / 2).numpy(), tab16, src.encoder.offsets.cpu().numpy(), S, 16)
    y, _ = O.mlp_forward(feat, w, 32, 64, 2)
    np.exp(y[:, 0].astype(np.float32))
    cpu_s = time.time() - tc
    params = 64 * (32 + 64 + 16)
    nbytes = M * (12 + 512 + 32 + 4)
    line = {"metric": "fused encoder->sigma-MLP inference rays/sec (device-timed), 4096 rays x 1024 samples", "value": NR / (ms * 1e-3), "unit": "rays/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "c3_fused_density_inference_4096x1024", "points": M, "caller": who,
                       "l2": "flushed before every timed iteration; 50 MB of points + 134 MB of outputs per iteration"},
            "e2e": {"value": NR / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": M * 12, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
            "gpu_launches": nlaunch, "clocks": _clocks(sampler, t0, t1),
            "roofline": _roof("ngp_field_sigma_forward", k[0], nbytes, 2 * params * M, pk) if k else None,
            "kernel_time_share": (k[0] / ms) if k else None,
            "cpu_baseline": {"value": (n / 1024) / cpu_s, "unit": "rays/s", "cores": int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)),
                             "kind": "port", "sample": f"oracle grid forward + numpy MLP + exp on 64 rays x 1024 samples ({cpu_s:.2f} s)"},
            "ref_cuda": ref}
    _finish(line, args)


# ================================================================================================ C5
def _sdf_points(B, seed=7):
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(B, 3, generator=g); d = d / d.norm(dim=-1, keepdim=True)
    X = d * 0.8
    X[: B // 2] += torch.randn(B // 2, 3, generator=g) * 0.01          # half of the surface points perturbed (sdf/provider.py:66-73)
    X[B * 7 // 8:] = torch.rand(B - B * 7 // 8, 3, generator=g) * 2 - 1
    X = X.clamp(-1, 1)
    return X, X.norm(dim=-1, keepdim=True) - 0.8


def run_c5(args, dev, pk, sampler):
    from oracle import oracle as O, ref_stack
    B = 1 << 20
    X_host, Y_host = _sdf_points(B)
    X_host, Y_host = X_host.contiguous().pin_memory(), Y_host.contiguous().pin_memory()
    X, Y = X_host.to(dev), Y_host.to(dev)
    flush = _flush_buf(dev)

    def make(stack):
        sdf = stack.module("sdf.netowrk_ff")
        with stack.active():
            torch.manual_seed(2)
            net = sdf.SDFNetwork().to(dev).train()
            with torch.no_grad():
                net.encoder.embeddings.uniform_(-0.1, 0.1)
        return net

    def step_fn(stack, net):
        def f(x=X, y=Y):
            with stack.active():
                net.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.float16):
                    p = net(x)
                    loss = ((p - y).abs() / (y.abs() + 1e-2)).mean()         # loss.py mape_loss
                (loss * 1024.0).backward()
            return loss
        return f

    ours = ref_stack.load("ours")
    net = make(ours)
    f = step_fn(ours, net)
    t0 = sampler.mark()
    ms = _time_iters(f, args.steps, max(args.warmup, 3), flush)
    t1 = sampler.mark()
    calls, nlaunch = _profile_calls(f, flush)
    xs, ys = torch.empty_like(X), torch.empty_like(Y)

    def e2e():
        xs.copy_(X_host, non_blocking=True); ys.copy_(Y_host, non_blocking=True)
        return float(f(xs, ys).item())
    ms_e2e = _time_iters(e2e, max(3, args.steps // 2), 1, flush)
    ref = None
    if ref_stack.available("ref"):
        rs = ref_stack.load("ref")
        rnet = make(rs)
        rnet.load_state_dict(net.state_dict())
        fr = step_fn(rs, rnet)
        ms_ref = _time_iters(fr, max(3, min(args.steps, 10)), 2, flush)
        ref = {"value": B / (ms_ref * 1e-3), "unit": "points/s", "ms_per_iter": ms_ref, "loss_ref": float(fr()), "loss_ours": float(f()),
               "what": "reference sdf/netowrk_ff.py SDFNetwork over the reference's wrappers + CUDA extensions (oracle/_ref), same points, same GPU"}
    dom = max(calls, key=lambda k_: calls[k_][0])
    per = {"ngp_field_sigma_forward": (12 + 512 + 64 + 3 * 128 + 32, 2 * 64 * (32 + 128 + 16)),
           "ngp_ffmlp_backward": (32 + 64 + 3 * 128 + 64, 4 * 64 * (32 + 128 + 16)), "ngp_grid_encode_backward": (588, 0)}
    by, fl = per.get(dom, (588, 0))
    n = 1 << 15
    tab16 = net.encoder.embeddings.detach().half().cpu().numpy()
    S = float(np.log2(net.encoder.per_level_scale))
    w = net.backbone.weights.detach().half().cpu().numpy()
    tc = time.time()
    x01 = ((X_host[:n] + 1) / 2).numpy()
    feat = O.grid_forward(x01, tab16, net.encoder.offsets.cpu().numpy(), S, 16)
    yo, fw = O.mlp_forward(feat, w, 32, 64, 3)
    gy = np.zeros_like(yo); gy[:, 0] = 1.0
    gi, _gw = O.mlp_backward(gy, feat, w, fw, 32, 64, 3)[:2]
    O.grid_backward(np.asarray(gi, dtype=np.float16), x01, net.encoder.offsets.cpu().numpy(), tab16.shape[0], 2, S, 16)
    cpu_s = time.time() - tc
    line = {"metric": "SDF field fwd+bwd points/sec (device-timed), hashgrid + FFMLP(32-64-64-64-1), 2^20 synthetic surface points",
            "value": B / (ms * 1e-3), "unit": "points/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "c5_sdf_1M_points", "points": B, "caller": "reference sdf/netowrk_ff.py SDFNetwork (unmodified) over this repo's packages; mape loss; autograd",
                       "l2": "flushed before every timed iteration"},
            "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "points/s", "h2d_bytes_per_step": B * 16, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
            "gpu_launches": nlaunch, "clocks": _clocks(sampler, t0, t1),
            "roofline": _roof(dom, calls[dom][0], B * by, B * fl, pk),
            "kernels": {k_: {"ms": v[0], "calls": v[2]} for k_, v in sorted(calls.items(), key=lambda kv: -kv[1][0])},
            "kernel_time_share": sum(v[0] for v in calls.values()) / ms,
            "cpu_baseline": {"value": n / cpu_s, "unit": "points/s", "cores": int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)), "kind": "port",
                             "sample": f"oracle grid fwd + numpy MLP fwd/bwd + grid bwd on 32768 points ({cpu_s:.2f} s)"},
            "ref_cuda": ref}
    _finish(line, args)


# ================================================================================================ inference render
def run_infer(args, dev, pk, sampler):
    import bench
    from oracle import ref_stack
    src, _ = bench.build_model(dev)
    with torch.no_grad():
        src.encoder.embeddings.uniform_(-0.3, 0.3)
    host_in, dev_in = bench.make_inputs(800 * 800, 0, 1, dev)
    N = 800 * 800
    flush = _flush_buf(dev)

    def make(stack):
        m = ref_stack.make_nerf(stack, bound=1).to(dev).eval()
        m.load_state_dict(src.state_dict(), strict=False)
        return m

    def render_fn(stack, m):
        def f(ro=dev_in[0][0], rd=dev_in[0][1]):
            with stack.active(), torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                return m.render(ro[None], rd[None], staged=True, bg_color=1, perturb=False, dt_gamma=0, max_steps=1024)["image"]
        return f

    ours = ref_stack.load("ours")
    mo = make(ours)
    f = render_fn(ours, mo)
    t0 = sampler.mark()
    ms = _time_iters(f, args.steps, max(args.warmup, 3), flush)
    t1 = sampler.mark()
    calls, nlaunch = _profile_calls(f, flush)
    so, sd = torch.empty_like(dev_in[0][0]), torch.empty_like(dev_in[0][1])

    def e2e():
        so.copy_(host_in[0][0], non_blocking=True); sd.copy_(host_in[0][1], non_blocking=True)
        return f(so, sd).float().cpu()
    ms_e2e = _time_iters(e2e, max(3, args.steps // 2), 1, flush)
    # the same frame with the fused field driver of this repo (nerf_step.render_eval: fused field kernels, alive-ray compaction
    # on the device, one host read per 8 iterations)
    fused = None
    try:
        from nerf_step import render_eval
        src.eval()

        def ff():
            return render_eval(src, dev_in[0][0], dev_in[0][1], bg_color=1.0)["image"]
        ms_f = _time_iters(ff, args.steps, 2, flush)
        d = float((ff().float() - f()[0].float()).abs().max())
        fused = {"value": N / (ms_f * 1e-3), "unit": "rays/s", "ms_per_frame": ms_f, "max_abs_diff_vs_reference_caller": d,
                 "what": "nerf_step.render_eval: same march/composite kernels, fused field kernels, device-side alive-ray compaction"}
    except ImportError:
        pass
    ref = None
    if ref_stack.available("ref"):
        rs = ref_stack.load("ref")
        mr = make(rs)
        fr = render_fn(rs, mr)
        ms_ref = _time_iters(fr, max(3, min(args.steps, 10)), 2, flush)
        ref = {"value": N / (ms_ref * 1e-3), "unit": "rays/s", "ms_per_frame": ms_ref,
               "max_abs_diff_vs_ours": float((fr().float() - f().float()).abs().max()),
               "what": "reference renderer + NeRFNetwork over the reference's wrappers + CUDA extensions (oracle/_ref), same frame, same GPU"}
    dom = max(calls, key=lambda k_: calls[k_][0])
    line = {"metric": "inference rays/sec (device-timed), full 800x800 frame through run_cuda's eval branch", "value": N / (ms * 1e-3), "unit": "rays/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "infer_800x800_synthetic_lego_boxes", "rays": N, "caller": "reference nerf/renderer.py run_cuda eval branch + nerf/network_ff.py (unmodified) over this repo's packages",
                       "l2": "flushed before every timed frame"},
            "e2e": {"value": N / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": N * 24, "d2h_bytes_per_step": N * 12, "ms_per_step": ms_e2e},
            "gpu_launches": nlaunch, "clocks": _clocks(sampler, t0, t1),
            "kernels": {k_: {"ms": v[0], "calls": v[2]} for k_, v in sorted(calls.items(), key=lambda kv: -kv[1][0])},
            "kernel_time_share": sum(v[0] for v in calls.values()) / ms, "dominant": dom,
            "fused_driver": fused, "ref_cuda": ref, "cpu_baseline": None}
    _finish(line, args)


def main(args):
    import bench
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    import _ngp_b200 as nb
    nb.load()
    pk = bench.peaks()
    sampler = bench.ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    {"c1": run_c1, "c3": run_c3, "c5": run_c5, "infer": run_infer}[args.config](args, dev, pk, sampler)
