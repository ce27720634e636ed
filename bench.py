#!/usr/bin/env python
"""bench.py — training rays/sec of the fused NeRF hot path on synthetic 800x800 rays (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W        # CPU arm (pure-PyTorch fp32 path, oracle/torch_ref.py)

One "step" = one optimisation step over `--rays-per-step` rays (default: a full 800x800 frame = 640 000 rays) of a
synthetic blender-format scene: near/far -> occupancy-grid march (steady-state mean_count path) -> hash-grid encode ->
sigma MLP -> SH -> color MLP -> composite -> MSE -> backward (-> gradient allreduce at N>1) -> GradScaler + Adam.
Prints ONE JSON line (rank 0).  `value` = device-timed whole-job rays/s with inputs resident in HBM; `e2e` = the same
through the public API with per-step pinned-host -> device copies of the rays/targets and a device -> host read of
the loss inside the timed region.  Strong scaling: the step's rays are split across ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200")]

import numpy as np   # noqa: E402
import torch         # noqa: E402

H_IMG = W_IMG = 800
METRIC = "training rays/sec (device-timed) at 800x800, L=16 hashgrid"
UNIT = "rays/s"
N_CAMERAS = 4          # distinct full frames cycled through (inputs >> L2: ~3-8 GB of activations per step)


_T0 = time.time()


def log(msg):
    print(f"[bench +{time.time() - _T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays-per-step", type=int, default=H_IMG * W_IMG)
    ap.add_argument("--cpu-rays", type=int, default=512, help="rays per CPU-baseline step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-maintenance", action="store_true", help="leave the occupancy-grid update (every 16th step) out of the timed region")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the steady-state step in a CUDA graph")
    ap.add_argument("--torch-optimizer", action="store_true", help="GradScaler + torch fused Adam on fp32 .grad (reference trainer sequence) instead of the fused fp16-sink optimizer kernel")
    ap.add_argument("--unfused", action="store_true", help="evaluate the field module by module (network_ff.py call sequence) instead of the fused field kernels")
    ap.add_argument("--chunks", type=int, default=1, help="row chunks of the fused field (side-stream pipelining of color fwd / table scatter); 1 = off")
    ap.add_argument("--no-prefetch", action="store_true", help="march each step's rays inside that step instead of one step ahead on a low-priority stream")
    ap.add_argument("--prefetch-point", default="auto", choices=["auto", "start", "exchange"])
    ap.add_argument("--mlp-backward", default="dual", choices=["dual", "single"], help="two-context MLP backward kernel (default) or the single-context one")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"], help="gradient exchange at N > 1: fused peer-memory reduce-scatter + sharded Adam + operand all-gather (csrc/exchange.cu; auto = use it when the ranks can map each other's memory) or NCCL all-reduce + full optimizer pass per rank")
    ap.add_argument("--long-steps", type=int, default=200, help="extra, longer timed region reported as `long_run` (0 = skip)")
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c5", "infer"], help="BASELINE.json config: c2 = training step (default; c4 = the same under torchrun), c1 = GridEncoder fwd/bwd 64k points, c3 = fused density inference 4096x1024, c5 = SDF 1M points, infer = full-frame eval render")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- helpers
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]), tf_sust=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows[-3:]]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(rows)}


def units_of(name, args):
    """(samples_or_rays, algorithmic bytes, flops) of one C-ABI call — SURVEY §8(d) per-unit figures."""
    if name == "ngp_grid_encode_forward":
        B, D, C, L, dtype = args[4], args[5], args[6], args[7], args[14]
        sz = 2 if dtype == 1 else 4
        return B, B * (4 * D + L * (1 << D) * C * sz + L * C * sz), 0
    if name == "ngp_grid_encode_backward":
        B, D, C, L, dtype = args[5], args[6], args[7], args[8], args[16]
        sz = 2 if dtype == 1 else 4
        return B, B * (4 * D + L * C * sz + L * (1 << D) * C * sz), 0
    if name in ("ngp_ffmlp_forward", "ngp_ffmlp_inference"):
        B, ind, outd, hid, nl = args[2], args[3], args[4], args[5], args[6]
        params = hid * (ind + hid * (nl - 1) + outd)
        return B, B * 2 * (ind + outd + (nl * hid if name == "ngp_ffmlp_forward" else 0)), 2 * params * B
    if name in ("ngp_ffmlp_backward", "ngp_ffmlp_backward_ex"):
        B, ind, outd, hid, nl = args[4], args[5], args[6], args[7], args[8]
        params = hid * (ind + hid * (nl - 1) + outd)
        return B, B * 2 * (outd + ind + nl * hid + ind), 4 * params * B      # reads dY, X, forward stash; writes dX
    if name == "ngp_field_sigma_forward":
        L, nl, M, train = args[3], args[9], args[10], args[11]
        params = 64 * (2 * L + 64 * (nl - 1) + 16)
        # 12 B xyz + L*8 corners*2 features*2 B gathers (512 B at L=16) + [feature stash 4L + hidden stash nl*128] + h 32 B + sigma 4 B
        return M, M * (12 + L * 8 * 2 * 2 + (2 * 2 * L + nl * 128 if train else 0) + 32 + 4), 2 * params * M
    if name == "ngp_field_color_forward":
        nl, M, train = args[3], args[4], args[5]
        params = 64 * (32 + 64 * (nl - 1) + 16)
        return M, M * (12 + 32 + (nl * 128 if train else 0) + 12), 2 * params * M
    if name in ("ngp_field_color_backward", "ngp_field_color_backward_ex"):
        nl, M = (args[7], args[8]) if name == "ngp_field_color_backward" else (args[9], args[10])
        params = 64 * (32 + 64 * (nl - 1) + 16)
        return M, M * (12 + 12 + 4 + 32 + 12 + nl * 128 + 32), 4 * params * M
    if name == "ngp_march_rays_train":
        return args[6], None, 0      # bytes depend on the emitted sample count (filled in by the caller)
    if name == "ngp_composite_rays_train_forward":
        M, N = args[4], args[5]
        return N, 24 * M + 32 * N, 0
    if name == "ngp_composite_rays_train_forward_mse":
        M, N = args[4], args[5]
        return N, 24 * M + (32 + 12 + 12 + 4 + 4) * N, 0         # + target read, g_image / g_ws / sqerr written
    if name == "ngp_exchange_reduce_fused":
        world, count = args[4], args[6]
        return count, count * 2 * (world + 1), 0                  # my shard read from every rank's bucket, reduced copy written
    if name == "ngp_composite_rays_train_backward":
        M, N = args[8], args[9]
        return N, 40 * M + 44 * N, 0
    if name == "ngp_sh_encode_forward":
        B, deg = args[2], args[4]
        return B, B * (12 + 4 * deg * deg), 0
    if name == "ngp_near_far_from_aabb":
        return args[3], args[3] * 32, 0
    return 0, 0, 0


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_arm(args, rays_per_step, steps, warmup):
    from oracle import torch_ref as T
    import ngp_synth as S
    cores = min(os.cpu_count() or 1, 32)     # torch CPU kernels stop scaling (and start thrashing) far below a 200-thread host
    torch.set_num_threads(cores)
    poses = S.make_cameras(N_CAMERAS, seed=11)
    g = torch.Generator().manual_seed(3)
    n_pool = max(rays_per_step * 4, 2048)
    inds = torch.randint(0, H_IMG * W_IMG, (n_pool,), generator=g)
    ro, rd = S.get_rays(poses[0], S.intrinsics(H_IMG, W_IMG), H_IMG, W_IMG, inds)
    target = torch.rand(n_pool, 3, generator=g)
    sec, loss = T.train_steps(rays_per_step, steps, warmup, ro, rd, target, threads=cores)
    return dict(value=rays_per_step / sec, unit=UNIT, cores=cores, kind="port",
                sample=f"{steps} steps x {rays_per_step} random rays of the 800x800 frame, 512 samples/ray (reference `run` path), "
                       f"fp32 torch on {cores} host threads; {sec * 1e3:.1f} ms/step"), sec


# ----------------------------------------------------------------------------------------------- GPU arm
def build_model(dev, fused=True):
    import ngp_synth as S
    from nerf_step import NeRFFieldFF
    torch.manual_seed(1)
    model = NeRFFieldFF(bound=1, fused=fused).to(dev).train()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        model.encoder.embeddings.copy_((torch.rand(model.encoder.embeddings.shape, generator=g) * 2 - 1) * 1e-4)
    grid, fill = S.box_union_density(128, seed=12)
    model.density_bitfield.copy_(torch.from_numpy(S.packbits_np(grid.numpy())).to(dev))
    return model, fill


def make_inputs(rays_total, rank, world, dev):
    """Per-camera (rays_o, rays_d, target) for this rank's ray range, pinned on the host and resident on the device."""
    import ngp_synth as S
    import ngp_dp
    poses = S.make_cameras(N_CAMERAS, seed=11)
    intr = S.intrinsics(H_IMG, W_IMG)
    host, devs = [], []
    for c in range(N_CAMERAS):
        g = torch.Generator().manual_seed(100 + c)
        if rays_total == H_IMG * W_IMG:
            inds = torch.arange(H_IMG * W_IMG)
        else:
            inds = torch.randint(0, H_IMG * W_IMG, (rays_total,), generator=g)
        mine = ngp_dp.shard_indices(rays_total, rank, world)
        ro, rd = S.get_rays(poses[c], intr, H_IMG, W_IMG, inds[mine])
        tgt = torch.rand(rays_total, 3, generator=g)[mine].contiguous()
        h = tuple(t.contiguous().pin_memory() for t in (ro, rd, tgt))
        host.append(h)
        devs.append(tuple(t.to(dev) for t in h))
    return host, devs


_REAL_STDOUT = None


def emit(line):
    """The bench contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner to fd 1) must not add
    to it: fd 1 is pointed at stderr for the duration of the run and the line is written to the saved descriptor."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    import faulthandler
    faulthandler.dump_traceback_later(240, repeat=True, file=sys.stderr)
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        # CPU arm: rank 0 alone runs; other ranks exit 0 without work
        if rank != 0:
            return
        cb, sec = cpu_arm(args, args.cpu_rays, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "nerf_train_800x800_synthetic_lego_boxes", "rays_per_step": args.rays_per_step,
                           "hashgrid": "L=16 F=2 T=2^19 base16 ->2048",
                           "note": f"reference pure-PyTorch path (--fp32, no --cuda_ray; renderer.py:125-253 `run`, 512 samples/ray) on the host cores; each step is a "
                                   f"bounded sample of {args.cpu_rays} rays of the same 800x800 frame (cpu_baseline.sample)"},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: CUDA device required (the product path has no CPU fallback); use --impl reference for the CPU arm")
    if args.config != "c2":
        # the other BASELINE.json configs: single-GPU op / inference throughput arms (bench_configs.py); one JSON line each
        if rank == 0:
            import bench_configs
            bench_configs.main(args)
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import _ngp_b200 as nb
    import ngp_dp
    nb.load().ngp_debug_set_mlp_backward(1 if args.mlp_backward == "dual" else 0)

    R = args.rays_per_step
    n_local = int(ngp_dp.shard_indices(R, rank, world).numel())
    log("building model")
    model, fill = build_model(dev, fused=not args.unfused)
    if world > 1:
        ngp_dp.broadcast_module(model)
    log("making inputs")
    host_in, dev_in = make_inputs(R, rank, world, dev)
    log("inputs ready")
    params = [model.encoder.embeddings, model.sigma_net.weights, model.color_net.weights]
    use_fused_opt = not (args.torch_optimizer or args.unfused)
    # static device staging: rays of the NEXT step (marched one step ahead) and targets of the CURRENT step
    stage = tuple(torch.empty_like(t) for t in dev_in[0])
    prefetch = use_fused_opt and not args.no_prefetch
    fstep = None
    if use_fused_opt:
        from ngp_optim import FusedFieldOptimizer
        fopt = FusedFieldOptimizer(model.encoder, model.sigma_net, model.color_net, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, init_scale=128.0,
                                   exchange=args.exchange)
        log(f"gradient exchange: {('peer memory (fused with the optimizer) on ' + str(fopt.px.memory)) if fopt.px is not None else ('nccl all-reduce' if world > 1 else 'none (1 GPU)')}"
            + (f" [symmetric memory unavailable: {fopt.px._symm_error}]" if fopt.px is not None and getattr(fopt.px, '_symm_error', None) else ""))

        from nerf_step import FusedTrainStep
        # where the next step's march is released: "start" (under this step's forward kernels, which leave registers / shared memory for
        # its 64-thread blocks) measured faster than "exchange" at every N (8 GPUs, 200 steps: 1.145 vs 1.19 ms; profiles/r2_scale8_*.json)
        ppoint = args.prefetch_point if args.prefetch_point != "auto" else "start"
        fstep = FusedTrainStep(model, fopt, R, perturb=True, chunks=args.chunks, prefetch_point=ppoint)

        def step(ro, rd, tgt):
            """budget-establishing step (no sample budget yet): autograd through the fused field, same optimizer"""
            with torch.autocast("cuda", dtype=torch.float16):
                out = model.render_train(ro, rd, perturb=True)
                loss = ((out["image"] - tgt) ** 2).sum() / (3.0 * R)
            (loss * fopt.scale_tensor()).backward()      # fp16 grads land in the flat sink (no fp32 .grad)
            fopt.step()                                  # [allreduce fp16 sink] + inf check + Adam + shadow refresh + zero
            return loss, out
    else:
        bucket = ngp_dp.FlatGradBucket(params)
        opt = torch.optim.Adam(params, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, fused=True)
        scaler = torch.amp.GradScaler("cuda", init_scale=128.0)

        def step(ro, rd, tgt):
            bucket.zero()
            with torch.autocast("cuda", dtype=torch.float16):
                out = model.render_train(ro, rd, perturb=True)
                # sum over local rays / global ray count: averaging the allreduce over ranks is then not needed
                loss = ((out["image"] - tgt) ** 2).sum() / (3.0 * R)
            scaler.scale(loss).backward()
            if world > 1:
                bucket.allreduce(average=False)
            scaler.step(opt)
            scaler.update()
            return loss, out

    # ---- establish the steady-state sample budget (reference: mean_count after the first epoch) ----
    counts = []
    for c in range(N_CAMERAS):
        model.mean_count = 0
        loss, out = step(*dev_in[c])
        counts.append(int(model.step_counter[(model.local_step - 1) % 16, 0].item()))
        log(f"budget step {c}: samples={counts[-1]} loss={float(loss.detach()):.5f}")
    model.mean_count = max(counts)       # no ray is dropped in the timed region (dropping = skipped work)
    samples_per_step_local = float(np.mean(counts))

    # ---- the steady-state step.  Fused path: FusedTrainStep (no autograd, no host sync, device-side sample budget / loss scale /
    # optimizer).  Step i consumes the samples marched during step i-1 and marches the rays of step i+1 on a low-priority stream
    # (`prefetch`), so one call = {march(i+1) || [forward(i), backward(i), exchange(i), optimizer(i)]}.  The whole step runs on a
    # high-priority stream so that the prefetch only fills idle issue slots.  Two CUDA graphs (even / odd sample slot) are captured
    # and replayed alternately; inputs are copied into the graphs' static buffers.
    it_global = [0]          # step index, continued across the warm-up / timed / e2e loops (camera i % N_CAMERAS)
    hp = torch.cuda.Stream(priority=-1)

    def fused_call(tgt, nro, nrd):
        if prefetch:
            return fstep.step_prefetched(tgt, nro, nrd)
        return fstep(nro, nrd, tgt)

    graphs = None
    graph_loss = [None, None]
    if use_fused_opt:
        # first samples: camera 0 (consumed by the first steady-state step)
        hp.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(hp):
            if prefetch:
                fstep.march(dev_in[0][0], dev_in[0][1])
            for i in range(4):      # warm the allocator pools / both slots eagerly
                c, cn = it_global[0] % N_CAMERAS, (it_global[0] + 1) % N_CAMERAS
                if prefetch:
                    fused_call(dev_in[c][2], dev_in[cn][0], dev_in[cn][1])
                else:
                    fused_call(dev_in[c][2], dev_in[c][0], dev_in[c][1])
                it_global[0] += 1
        torch.cuda.current_stream().wait_stream(hp)
        torch.cuda.synchronize()
    if use_fused_opt and not args.no_graph:
        try:
            graphs = []
            for par in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=hp):
                    graph_loss[par] = fused_call(stage[2], stage[0], stage[1])
                graphs.append(g)
                torch.cuda.synchronize()
            log("CUDA graphs captured (even / odd sample slot)")
        except Exception as e:
            log(f"CUDA graph capture failed, running eagerly: {type(e).__name__}: {str(e)[:300]}")
            graphs = None
            torch.cuda.synchronize()

    # ---- occupancy-grid maintenance: the reference trainer calls update_extra_state every 16 steps
    # (nerf/utils.py train_one_epoch).  Its cost belongs in the training throughput, but its RESULT must not replace the synthetic
    # scene's fixed occupancy (a random-init field marks every cell occupied, which would change the workload), so it runs on a
    # shadow copy of the grid: same kernels, same sample counts (steady-state "partial" update: 2 x H^3/4 samples per cascade),
    # same host read, result discarded.
    maint = None
    if not args.no_maintenance and not args.unfused:
        import types
        import density_grid as dg
        import ngp_synth as S
        g0, _ = S.box_union_density(128, seed=12)
        shadow = types.SimpleNamespace(cuda_ray=True, density_grid=g0.to(dev).contiguous(), density_bitfield=torch.zeros_like(model.density_bitfield),
                                       cascade=model.cascade, grid_size=model.grid_size, bound=model.bound, density_scale=model.density_scale,
                                       density_thresh=model.density_thresh, iter_density=16, mean_density=0.0, mean_count=0, local_step=0,
                                       step_counter=model.step_counter, encoder=model.encoder, sigma_net=model.sigma_net,
                                       color_net=model.color_net)

        def maint():
            shadow.iter_density = 16
            dg.update_extra_state(shadow)      # rank-local; replicas would be re-synchronised with ngp_dp.sync_occupancy (one 8 MB broadcast per 16
                                               # steps, ~10 us over NVLink) or kept identical with ngp_dp.seed_lock — no collective is issued here
        try:
            maint()
            torch.cuda.synchronize()
        except Exception as e:      # measured without maintenance rather than not at all; the JSON line says which
            log(f"occupancy maintenance unavailable, left out of the timed region: {type(e).__name__}: {str(e)[:200]}")
            maint = None

    copy_stream = torch.cuda.Stream()
    landing = [tuple(torch.empty_like(t) for t in dev_in[0]) for _ in range(2)]
    copy_done = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    for ev in consumed:
        ev.record()

    def host_triplet(i):
        """(rays_o, rays_d) of the step whose samples are generated during step i, and the target step i consumes."""
        c = i % N_CAMERAS
        cn = (i + 1) % N_CAMERAS if prefetch else c
        return host_in[cn][0], host_in[cn][1], host_in[c][2]

    def dev_triplet(i):
        c = i % N_CAMERAS
        cn = (i + 1) % N_CAMERAS if prefetch else c
        return dev_in[cn][0], dev_in[cn][1], dev_in[c][2]

    def run_loop(n, e2e, eager=False):
        last = None
        cur = torch.cuda.current_stream()
        for k in range(n):
            i = it_global[0]
            if use_fused_opt:
                if e2e:
                    # pinned host -> device on a copy stream, one step ahead (double-buffered landing zones), so the PCIe
                    # transfer of step i+1 overlaps the compute of step i; every step's bytes are still copied inside the
                    # timed region
                    if k == 0:
                        with torch.cuda.stream(copy_stream):
                            copy_stream.wait_event(consumed[i % 2])
                            for dst, src in zip(landing[i % 2], host_triplet(i)):
                                dst.copy_(src, non_blocking=True)
                            copy_done[i % 2].record(copy_stream)
                    if k + 1 < n:
                        with torch.cuda.stream(copy_stream):
                            copy_stream.wait_event(consumed[(i + 1) % 2])
                            for dst, src in zip(landing[(i + 1) % 2], host_triplet(i + 1)):
                                dst.copy_(src, non_blocking=True)
                            copy_done[(i + 1) % 2].record(copy_stream)
                    cur.wait_event(copy_done[i % 2])
                    src3 = landing[i % 2]
                else:
                    src3 = dev_triplet(i)
                if graphs is not None and not eager:
                    for dst, src in zip(stage, src3):
                        dst.copy_(src, non_blocking=True)
                    if e2e:
                        consumed[i % 2].record()
                    graphs[i % 2].replay()
                    if e2e:
                        last = graph_loss[i % 2].item()                # device -> host read of the step's result
                else:
                    fstep.set_slot(i)
                    hp.wait_stream(cur)
                    with torch.cuda.stream(hp):
                        loss = fused_call(src3[2], src3[0], src3[1])
                    cur.wait_stream(hp)
                    if e2e:
                        consumed[i % 2].record()
                        last = loss.item()
            elif e2e:
                c = i % N_CAMERAS
                for dst, src in zip(stage, host_in[c]):
                    dst.copy_(src, non_blocking=True)
                loss, _ = step(*stage)
                last = loss.item()            # device -> host read of the step's result
            else:
                loss, _ = step(*dev_in[i % N_CAMERAS])
            it_global[0] += 1
            if maint is not None and (k + 1) % 16 == 0:
                maint()
        return last

    def timed(n, e2e, profile=False, eager=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        if profile:
            nb.profile_begin()
        nb.reset_launch_count()
        e0.record()
        run_loop(n, e2e, eager=eager)
        e1.record()
        torch.cuda.synchronize()
        launches = nb.launch_count()
        rec = nb.profile_end() if profile else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), launches, rec

    W = max(args.warmup, 3)
    run_loop(W, False)            # (step index i <-> sample slot i % 2 <-> graph i % 2 throughout)
    torch.cuda.synchronize()
    log("warm-up done")
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    t0 = sampler.mark()
    long_run = None
    if use_fused_opt:
        ms_total, _, _ = timed(args.steps, False)                                      # the headline timing (graph replay unless --no-graph)
        t1 = sampler.mark()
        if args.long_steps > args.steps:
            ms_long, _, _ = timed(args.long_steps, False)
            long_run = {"steps": args.long_steps, "ms_per_step": ms_long / args.long_steps, "value": R * args.long_steps / (ms_long * 1e-3),
                        "timed_region_s": ms_long * 1e-3}
            t1 = sampler.mark()
        clocks = sampler.stop(t0, t1)
        # per-kernel CUDA-event durations need individual, non-overlapping launches: the same K steps once more, eagerly,
        # with the pipelining switched off (chunks = 1, march inside the step)
        saved = (fstep.chunks, prefetch)
        torch.cuda.synchronize()
        fstep.chunks, prefetch = 1, False
        run_loop(2, False, eager=True)      # re-warm the eager allocator pool (the graphs own theirs)
        ms_eager, launches, rec = timed(args.steps, False, profile=True, eager=True)
        log(f"eager unpipelined (per-kernel event pass): {ms_eager / args.steps:.2f} ms/step")
        fstep.chunks, prefetch = saved
        if prefetch:
            fstep.set_slot(it_global[0])
            fstep.march(*dev_triplet(it_global[0] - 1)[:2])     # refill the current slot for the loops that follow
            torch.cuda.synchronize()
    else:
        ms_total, launches, rec = timed(args.steps, False, profile=True)
        ms_eager = ms_total
        t1 = sampler.mark()
        clocks = sampler.stop(t0, t1)
    log(f"timed region done: {ms_total / args.steps:.2f} ms/step")
    run_loop(2, True)
    ms_e2e, _, _ = timed(args.steps, True)
    log(f"e2e done: {ms_e2e / args.steps:.2f} ms/step")

    value = R * args.steps / (ms_total * 1e-3)
    e2e_value = R * args.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host_in[0])

    # ---- per-entry-point device time inside the timed region -> dominant kernel + roofline ----
    agg = {}
    for name, a, s0, s1 in rec:
        n, by, fl = units_of(name, a)
        d = agg.setdefault(name, dict(ms=0.0, calls=0, bytes=0, flops=0, units=0))
        d["ms"] += s0.elapsed_time(s1); d["calls"] += 1; d["units"] += n
        if name == "ngp_march_rays_train":
            by = int(32 * samples_per_step_local + 48 * n)
        d["bytes"] += by or 0; d["flops"] += fl
    pk = peaks()
    kern_ms = sum(d["ms"] for d in agg.values())
    dom = max(agg, key=lambda k: agg[k]["ms"])
    dd = agg[dom]
    per_launch_ms = dd["ms"] / dd["calls"]
    ach_tf = dd["flops"] / dd["calls"] / (per_launch_ms * 1e-3) / 1e12
    ach_gb = dd["bytes"] / dd["calls"] / (per_launch_ms * 1e-3) / 1e9
    # a 64-wide MLP is 20-40 FLOP/B — far left of the B200 ridge (~220 FLOP/B): the binding roof is memory; the tensor
    # fraction is reported beside it
    if ach_tf / pk["tf_sust"] > ach_gb / pk["hbm"]:
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach_tf / pk["tf_sust"]}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach_gb, "peak": pk["hbm"], "unit": "GB/s", "frac": ach_gb / pk["hbm"],
                "tensor_tflops": ach_tf or None}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath)).get(dom)
        if t:   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, scaled per sample to this launch
            traffic = t["dram_bytes_per_sample"] * dd["units"] / dd["calls"]
    roof.update({"traffic": traffic, "traffic_source": (os.path.relpath(tpath, ROOT) + " (ncu --set full, per-sample x samples/launch)") if traffic else None,
                 "peak_source": pk["src"], "avg_launch_ms": per_launch_ms,
                 "share_of_step": dd["ms"] / ms_eager, "algorithmic_bytes_per_launch": dd["bytes"] / dd["calls"],
                 "units_per_launch": dd["units"] / dd["calls"],
                 "units_note": "sample rows of the padded steady-state budget M (mean_count rounded up to 128; pad rows are zero-filled and still gather / scatter), not only the rows a ray produced"})
    launches = launches // 1      # launches of OUR kernels per timed region (counted in the eager pass; a graph replays the same set)
    breakdown = {k: {"ms_per_step": v["ms"] / args.steps, "calls_per_step": v["calls"] / args.steps,
                     "GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["bytes"] and v["ms"] else None,
                     "TFLOPs": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] and v["ms"] else None} for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": "nerf_train_800x800_synthetic_lego_boxes", "rays_per_step": R, "rays_per_rank": n_local,
                       "samples_per_ray_mean": samples_per_step_local / max(1, n_local), "occupancy_fill": fill,
                       "hashgrid": "L=16 F=2 T=2^19 base16 ->2048", "mlp": "FFMLP 32-64-64-16 + 32-64-64-64-16 fp16/fp32-acc",
                       "optimizer": ("fused fp16-sink Adam kernel with device-side loss scaling (in timed region)" if use_fused_opt else "GradScaler + torch fused Adam (in timed region)"), "field_path": "module-by-module (network_ff.py sequence)" if args.unfused else "fused field kernels (nerf_fused.fused_field)", "parallelism": f"dp{world} (rays sharded in round-robin blocks of 256, 1 allreduce/step)",
                       "maintenance": ("occupancy-grid update_extra_state (partial update, 2 x 524288 samples/cascade, host read included) after every 16th step "
                                       "inside the timed region, on a shadow copy of the grid (the synthetic scene's occupancy is fixed)") if maint is not None else "none",
                       "l2": "inputs_exceed_l2 (per-step activations of several GB; 4 camera frames cycled)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world if world > 1 else h2d, "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "kernel_time_share": kern_ms / ms_eager, "kernels": breakdown,
            "cuda_graph": graphs is not None, "ms_per_step_eager_unpipelined": ms_eager / args.steps,
            "mlp_backward_kernel": args.mlp_backward,
            "exchange": ((f"peer-memory reduce-scatter + sharded Adam + fp16 operand all-gather in 3 launches with in-kernel flag barriers (csrc/exchange.cu); memory: {fopt.px.memory}; "
                          + ("multimem.ld_reduce / multimem.st through the NVSwitch" if fopt.px.mc_sink else "peer loads / stores")
                          if (use_fused_opt and fopt.px is not None) else "nccl all-reduce of the fp16 sink + full optimizer pass per rank") if world > 1 else "none"),
            "pipelining": {"field_chunks": args.chunks if use_fused_opt else 1, "march_prefetch": bool(prefetch),
                           "prefetch_point": (fstep.prefetch_point if fstep is not None else None),
                           "note": "per-kernel figures (`kernels`, `roofline`) come from an eager pass with the pipelining switched off, so that each launch is timed alone; the headline is the pipelined, graph-replayed step"},
            "long_run": long_run, "timed_region_s": ms_total * 1e-3}

    if use_fused_opt and fopt.px is not None:
        line["exchange_error"] = fopt.px.error()        # non-zero = a flag barrier timed out (the numbers above are then void)
        if line["exchange_error"]:
            log(f"PEER EXCHANGE ERROR {line['exchange_error']}")

    # ---- the table scatter is an L2-reduction kernel (the fp16 gradient table is L2-resident: dram traffic is 0.18x the algorithmic
    # bytes): report it against the device's MEASURED reduction rate for the same access pattern as well (random f16x2 / v2.f16x2
    # reductions into a 24.5 MB table, ngp_debug_red_probe), next to the HBM fraction the contract asks for
    try:
        n_entries = int(model.encoder.embeddings.shape[0]) // 4 * 4
        probe_tab = torch.zeros(n_entries, 2, dtype=torch.half, device=dev)
        blocks, ops = 148 * 16, 128
        rates = {}
        for mode, nm in ((0, "f16x2_4B"), (1, "v2_f16x2_8B")):
            for rep in range(2):
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record()
                nb.call("ngp_debug_red_probe", probe_tab.data_ptr(), n_entries, blocks, ops, mode)
                b.record()
                torch.cuda.synchronize()
            rates[nm] = blocks * 256 * ops / (a.elapsed_time(b) * 1e-3) / 1e9
        red_per_sample = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get("ngp_grid_encode_backward", {}).get("red_ops_per_sample") \
            if os.path.exists(os.path.join(ROOT, "profiles", "r2_traffic.json")) else None
        gb = agg.get("ngp_grid_encode_backward")
        line_red = {"peak_gops_measured": rates, "how": "ngp_debug_red_probe: 148*16 blocks x 256 threads x 128 random-entry reductions into a 24.5 MB fp16x2 table"}
        if red_per_sample and gb:
            ops_per_launch = red_per_sample * gb["units"] / gb["calls"]
            ach = ops_per_launch / (gb["ms"] / gb["calls"] * 1e-3) / 1e9
            line_red.update({"kernel": "ngp_grid_encode_backward", "red_ops_per_sample": red_per_sample, "achieved_gops": ach,
                             "frac_of_measured_8B_rate": ach / rates["v2_f16x2_8B"],
                             "source": "red ops per sample = lts__t_sectors_srcunit_tex_op_red.sum / rows of the ncu --set full capture (profiles/r2c_ncu_step.md)"})
        line["l2_reduction_roofline"] = line_red
        del probe_tab
    except Exception as e:
        line["l2_reduction_roofline"] = {"unavailable": str(e)[:200]}

    # side arms run in child processes with hard timeouts: a reported baseline must never cost the bench line
    def child(cmd, timeout):
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
            for ln in reversed(p.stdout.strip().splitlines()):
                if ln.startswith("{"):
                    return json.loads(ln)
            return {"unavailable": (p.stderr or "no output")[-300:]}
        except subprocess.TimeoutExpired:
            return {"unavailable": f"timed out after {timeout}s"}
        except Exception as e:
            return {"unavailable": str(e)[:300]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        log("cpu baseline (child process)")
        r = child([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                   "--cpu-rays", str(args.cpu_rays)], 240)
        line["cpu_baseline"] = r.get("cpu_baseline", {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                                      "sample": "failed: " + str(r.get("unavailable"))})
    if rank == 0 and world == 1 and not args.no_ref_cuda and not args.unfused:
        # the unchanged-caller path: the reference's own nerf/network_ff.py + nerf/renderer.py (unmodified, staged under oracle/_ref/py)
        # over this repo's four drop-in packages, autograd, GradScaler + torch Adam — what a torch-ngp user gets by swapping the
        # package directories and nothing else.  Deferred tensors (ngp_lazy) route the module sequence to the fused kernels.
        log("drop-in path: reference callers over our packages (child process)")
        torch.cuda.empty_cache()
        line["dropin"] = child([sys.executable, os.path.join(ROOT, "bench_ref_cuda.py"), "--stack", "ours", "--rays-per-step", str(R),
                                "--steps", "8"], 300)
        line["dropin_literal"] = child([sys.executable, os.path.join(ROOT, "bench_ref_cuda.py"), "--stack", "ours", "--no-lazy",
                                        "--rays-per-step", str(R), "--steps", "5"], 300)
    if rank == 0 and world == 1 and not args.no_ref_cuda:
        log("reference CUDA build arm (child process)")
        torch.cuda.empty_cache()
        line["ref_cuda"] = child([sys.executable, os.path.join(ROOT, "bench_ref_cuda.py"), "--stack", "ref", "--rays-per-step", str(R),
                                  "--steps", str(max(3, min(args.steps, 10)))], 300)
    if rank == 0:
        emit(line)
    # leave without library teardown: destroying a NCCL communicator that is referenced by a live CUDA graph can block
    torch.cuda.synchronize()
    if world > 1:
        try:
            dist.barrier()
        except Exception:
            pass
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
