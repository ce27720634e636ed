"""nerf_step.py — the caller side of the hot path, restated for measurement and tests.

The reference's callers (nerf/network_ff.py NeRFNetwork.forward/density :51-89 and the training branch of
nerf/renderer.py run_cuda :280-321) live in /root/reference and cannot travel to the GPU box, so this
module re-expresses exactly that call sequence on top of the drop-in packages
(gridencoder.GridEncoder -> ffmlp.FFMLP -> trunc_exp; shencoder.SHEncoder (+) geo_feat (+) pad ->
ffmlp.FFMLP -> sigmoid; raymarching.near_far_from_aabb / march_rays_train / composite_rays_train).
It is host glue, not a kernel; bench.py drives it, tests check it against the oracle.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from gridencoder import GridEncoder
from shencoder import SHEncoder
from ffmlp import FFMLP
import raymarching
import density_grid as _density_grid


class _trunc_exp(torch.autograd.Function):
    """activation.py:5-18 of the reference."""
    @staticmethod
    @torch.amp.custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @torch.amp.custom_bwd(device_type='cuda')
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _trunc_exp.apply


class NeRFFieldFF(nn.Module):
    """Field of nerf/network_ff.py:11-89 (hashgrid 16x2 -> FFMLP 32-64-64-16; SH4 (+) 15 geo (+) pad -> FFMLP 32-64-64-64-16)."""

    def __init__(self, bound=1, num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                 density_scale=1, min_near=0.2, density_thresh=0.01, grid_size=128, fused=False):
        super().__init__()
        self.fused = fused       # True: evaluate the field through nerf_fused.fused_field (same math, fused kernels)
        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.grid_size = grid_size
        self.density_scale = density_scale
        self.min_near = min_near
        self.density_thresh = density_thresh
        self.geo_feat_dim = geo_feat_dim
        self.encoder = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                                   desired_resolution=2048 * bound, gridtype='hash', align_corners=False)
        self.sigma_net = FFMLP(input_dim=self.encoder.output_dim, output_dim=1 + geo_feat_dim, hidden_dim=hidden_dim,
                               num_layers=num_layers)
        self.encoder_dir = SHEncoder(input_dim=3, degree=4)
        self.color_net = FFMLP(input_dim=self.encoder_dir.output_dim + geo_feat_dim + 1, output_dim=3,
                               hidden_dim=hidden_dim_color, num_layers=num_layers_color)
        aabb = torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound])
        self.register_buffer('aabb_train', aabb)
        self.cuda_ray = True
        self.register_buffer('density_grid', torch.zeros(self.cascade, grid_size ** 3))
        self.register_buffer('density_bitfield', torch.zeros(self.cascade * grid_size ** 3 // 8, dtype=torch.uint8))
        self.mean_density = 0
        self.iter_density = 0
        self.register_buffer('step_counter', torch.zeros(16, 2, dtype=torch.int32))
        self.mean_count = 0
        self.local_step = 0

    def forward(self, x, d):
        if self.fused:
            from nerf_fused import fused_field
            return fused_field(self.encoder, self.sigma_net, self.color_net, x, d, self.bound)
        x = self.encoder(x, bound=self.bound)
        h = self.sigma_net(x)
        sigma = trunc_exp(h[..., 0])
        geo_feat = h[..., 1:]
        d = self.encoder_dir(d)
        p = torch.zeros_like(geo_feat[..., :1])
        h = torch.cat([d, geo_feat, p], dim=-1)
        h = self.color_net(h)
        rgb = torch.sigmoid(h)
        return sigma, rgb

    def density(self, x):
        x = self.encoder(x, bound=self.bound)
        h = self.sigma_net(x)
        return {'sigma': trunc_exp(h[..., 0]), 'geo_feat': h[..., 1:]}

    def render_train(self, rays_o, rays_d, bg_color=1, perturb=True, force_all_rays=False, dt_gamma=0, max_steps=1024,
                     T_thresh=1e-4):
        """training branch of renderer.py run_cuda (:256-321)."""
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb_train, self.min_near)
        counter = self.step_counter[self.local_step % 16]
        counter.zero_()
        self.local_step += 1
        xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, self.bound, self.density_bitfield,
                                                                self.cascade, self.grid_size, nears, fars, counter,
                                                                self.mean_count, perturb, 128, force_all_rays, dt_gamma,
                                                                max_steps)
        sigmas, rgbs = self(xyzs, dirs)
        sigmas = self.density_scale * sigmas
        weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh)
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return {'image': image, 'depth': depth, 'weights_sum': weights_sum, 'n_samples': xyzs.shape[0]}

    # occupancy-grid maintenance: the fused kernels (density_grid.py) ...
    mark_untrained_grid = _density_grid.mark_untrained_grid
    update_extra_state = _density_grid.update_extra_state

    # ... and the reference's op-by-op torch sequence (renderer.py:445-530), restated for parity tests and timing
    @torch.no_grad()
    def update_extra_state_unfused(self, decay=0.95):
        H, dev = self.grid_size, self.density_bitfield.device
        fresh = torch.full_like(self.density_grid, -1)

        def query(cas, cells, slots):
            scale = min(2 ** cas, self.bound)
            half = scale / H
            pts = (2 * cells.float() / (H - 1) - 1) * (scale - half)
            pts += (torch.rand_like(pts) * 2 - 1) * half
            sig = self.density(pts)['sigma'].reshape(-1).detach().float()
            sig *= self.density_scale
            fresh[cas, slots] = sig

        if self.iter_density < 16:
            ax = torch.arange(H, dtype=torch.int32, device=dev)
            cells = torch.stack(torch.meshgrid(ax, ax, ax, indexing='ij'), dim=-1).reshape(-1, 3)
            slots = raymarching.morton3D(cells).long()
            for cas in range(self.cascade):
                query(cas, cells, slots)
        else:
            n = H ** 3 // 4
            for cas in range(self.cascade):
                cells = torch.randint(0, H, (n, 3), device=dev)
                slots = raymarching.morton3D(cells).long()
                occ = torch.nonzero(self.density_grid[cas] > 0).squeeze(-1)
                occ = occ[torch.randint(0, occ.shape[0], [n], dtype=torch.long, device=dev)]
                query(cas, torch.cat([cells, raymarching.morton3D_invert(occ)], dim=0), torch.cat([slots, occ], dim=0))
        both = (self.density_grid >= 0) & (fresh >= 0)
        self.density_grid[both] = torch.maximum(self.density_grid[both] * decay, fresh[both])
        self.mean_density = torch.mean(self.density_grid.clamp(min=0)).item()
        self.iter_density += 1
        self.density_bitfield = raymarching.packbits(self.density_grid, min(self.mean_density, self.density_thresh),
                                                     self.density_bitfield)
        self.update_mean_count()

    def update_mean_count(self):
        """tail of renderer.py update_extra_state (:532-536)."""
        total_step = min(16, self.local_step)
        if total_step > 0:
            self.mean_count = int(self.step_counter[:total_step, 0].sum().item() / total_step)
        self.local_step = 0


def train_step(model, rays_o, rays_d, target, optimizer=None, scaler=None, **render_kw):
    """One reference training iteration (nerf/utils.py:861-868): autocast forward, MSE, scaled backward, optimizer."""
    if optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    with torch.autocast('cuda', dtype=torch.float16):
        out = model.render_train(rays_o, rays_d, **render_kw)
        loss = torch.nn.functional.mse_loss(out['image'], target)
    if scaler is not None:
        scaler.scale(loss).backward()
        if optimizer is not None:
            scaler.step(optimizer)
            scaler.update()
    else:
        loss.backward()
        if optimizer is not None:
            optimizer.step()
    return loss.detach(), out


class FusedTrainStep:
    """Autograd-free, host-sync-free training step: near/far -> march (steady-state budget) -> fused field
    -> composite -> MSE -> composite backward -> fused field backward -> exchange -> fused optimizer, all issued from one thread
    through the C ABI, so the whole step can be captured in a CUDA graph and replayed.  Same arithmetic as train_step() with
    NeRFFieldFF(fused=True) + FusedFieldOptimizer: the MSE gradient is formed in closed form instead of by autograd.

    Three things overlap inside a step (all joined before it returns, so a captured graph is self-contained):
      * `chunks` > 1: the field is evaluated in row chunks; the color net of chunk k runs on a side stream under the encoder+sigma
        kernel of chunk k+1, and in the backward the hash-table scatter of chunk k runs under the MLP backward kernels of chunk k+1
        (nerf_fused.field_forward / field_backward);
      * `prefetch(rays_o, rays_d)` marches the NEXT step's rays on a low-priority stream while this step's backward, gradient
        exchange and optimizer run — the marcher is a latency-bound integer/ALU kernel that does not read the weights
        ("overlap must come from pipelining with the next step's R1/R5 march");
      * at world size > 1 the fp16 gradient sink is all-reduced asynchronously (NCCL's stream) while the prefetched march finishes.
    Samples live in two persistent (ping-pong) buffer sets sized for the model's steady-state budget `mean_count`."""

    def __init__(self, model, optimizer, rays_total, bg_color=1.0, T_thresh=1e-4, dt_gamma=0.0, max_steps=1024, perturb=True,
                 chunks=1, group=None, prefetch_point="start"):
        self.model, self.opt, self.R = model, optimizer, float(rays_total)
        self.bg, self.T_thresh, self.dt_gamma, self.max_steps, self.perturb = bg_color, T_thresh, dt_gamma, max_steps, perturb
        self.chunks, self.group = int(chunks), group
        assert prefetch_point in ("start", "exchange")
        self.prefetch_point = prefetch_point     # where the next step's march is released: with the forward, or with the exchange
        self._next = None                        # rays handed to step_prefetched(), marched inside body()
        self._pending = False
        self._slots = [None, None]      # marched sample sets (ping-pong)
        self._cur = 0                   # slot the next body() consumes
        self._ready = [False, False]
        self._side = None               # field pipelining stream
        self._march_stream = None       # prefetch stream (lowest priority)
        dev = model.density_bitfield.device
        self._ring = torch.zeros(1, dtype=torch.int32, device=dev)       # next row of model.step_counter (device side)
        self._nsteps = torch.zeros(1, dtype=torch.int32, device=dev)     # steps since the last sync_host_state()

    # ---- sample generation -------------------------------------------------------------------------------------------------
    def _streams(self):
        if self._side is None:
            # CUDA priorities: 0 is the LOWEST, negative is higher.  The prefetch march runs at the lowest priority; callers that want
            # it to yield to the step's own kernels run the step on a higher-priority stream (bench.py does).
            self._side = torch.cuda.Stream(priority=-1)
            self._march_stream = torch.cuda.Stream(priority=0)
        return self._side, self._march_stream

    def _slot(self, i, n_rays):
        m = self.model
        assert m.mean_count > 0, "FusedTrainStep runs the steady-state (mean_count) path; establish the budget first"
        M = m.mean_count + (128 - m.mean_count % 128)          # the reference's rule (raymarching.py:200-201)
        s = self._slots[i]
        if s is None or s["M"] != M or s["N"] != n_rays:
            dev = m.density_bitfield.device
            f = dict(dtype=torch.float32, device=dev)
            # xyzs | dirs | deltas | counter share one allocation: one memset per march instead of four
            flat = torch.zeros(M * 8 + 2, **f)
            s = dict(M=M, N=n_rays, flat=flat, xyzs=flat[:M * 3].view(M, 3), dirs=flat[M * 3:M * 6].view(M, 3),
                     deltas=flat[M * 6:M * 8].view(M, 2), counter=flat[M * 8:].view(torch.int32),
                     rays=torch.zeros(n_rays, 3, dtype=torch.int32, device=dev), nears=torch.empty(n_rays, **f),
                     fars=torch.empty(n_rays, **f),
                     # composite / loss scratch of the step that consumes these samples
                     grads=torch.zeros(M, 4, **f), per_ray=torch.empty(n_rays, 13, **f))
            self._slots[i] = s
        return s

    @torch.no_grad()
    def _march_into(self, i, rays_o, rays_d):
        """near/far + march of one ray batch into slot i, on the CURRENT stream."""
        import _ngp_b200 as nb
        m = self.model
        o = rays_o.contiguous().view(-1, 3)
        d = rays_d.contiguous().view(-1, 3)
        n = o.shape[0]
        s = self._slot(i, n)
        nb.call("ngp_near_far_from_aabb", o.data_ptr(), d.data_ptr(), m.aabb_train.data_ptr(), n, float(m.min_near),
                s["nears"].data_ptr(), s["fars"].data_ptr())
        # rows the marcher does not reach must read as zeros (raymarching.py:205-207)
        s["flat"].zero_()
        noise = torch.rand(n, dtype=torch.float32, device=o.device) if self.perturb else torch.zeros(n, dtype=torch.float32, device=o.device)
        nb.call("ngp_march_rays_train", o.data_ptr(), d.data_ptr(), m.density_bitfield.data_ptr(), float(m.bound), float(self.dt_gamma),
                int(self.max_steps), n, int(m.cascade), int(m.grid_size), s["M"], s["nears"].data_ptr(), s["fars"].data_ptr(),
                s["xyzs"].data_ptr(), s["dirs"].data_ptr(), s["deltas"].data_ptr(), s["rays"].data_ptr(), s["counter"].data_ptr(),
                noise.data_ptr())
        # the reference's 16-slot sample-count ring (renderer.py:281-283), advanced on the device
        nb.call("ngp_step_counter_push", self._ring.data_ptr(), s["counter"].data_ptr(), self._nsteps.data_ptr(), m.step_counter.data_ptr())
        self._ready[i] = True

    @torch.no_grad()
    def march(self, rays_o, rays_d):
        """Synchronous sample generation for the step that runs next."""
        self._march_into(self._cur, rays_o, rays_d)

    @torch.no_grad()
    def prefetch(self, rays_o, rays_d):
        """Start marching the FOLLOWING step's rays on the prefetch stream (ordered after the work queued so far on the current
        stream); body() joins it before returning."""
        _, ms = self._streams()
        ms.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(ms):
            self._march_into(1 - self._cur, rays_o, rays_d)
        self._pending = True

    def set_slot(self, i):
        """Re-synchronise the host-side slot bookkeeping after CUDA-graph replays (a replay runs no Python): step index i consumes
        slot i % 2, which the previous step (or march()) filled."""
        self._cur = int(i) % 2
        self._ready[self._cur] = True

    def sync_host_state(self):
        """One host read: hand the number of steps taken since the last call to the model (`local_step`), as the reference's
        update_extra_state expects before it re-estimates mean_count (renderer.py:532-536)."""
        n = int(self._nsteps.item())
        self._nsteps.zero_()
        self.model.local_step = self.model.local_step + n
        return n

    # ---- the step ----------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def body(self, target):
        """Forward, loss, backward, exchange and optimizer on the samples of the current slot; flips the slots afterwards."""
        import _ngp_b200 as nb
        from nerf_fused import field_forward, field_backward, field_cfg
        m = self.model
        s = self._slots[self._cur]
        assert s is not None and self._ready[self._cur], "no marched samples: call march() or prefetch() first"
        side, ms = self._streams()
        side_or_none = side if self.chunks > 1 else None
        if self._next is not None and self.prefetch_point == "start":
            self.prefetch(*self._next)
            self._next = None
        cfg = field_cfg(m.encoder, m.sigma_net, m.color_net, m.bound, True)
        sigma, rgb, stash = field_forward(s["xyzs"], s["dirs"], m.encoder.embeddings, m.encoder.offsets, m.sigma_net.weights,
                                          m.color_net.weights, cfg, chunks=self.chunks, side=side_or_none)
        if m.density_scale != 1:
            sigma = sigma * m.density_scale
        M, N = sigma.shape[0], s["N"]
        deltas, rays = s["deltas"], s["rays"]
        target = target.reshape(-1, 3).float().contiguous()
        assert target.shape[0] == N, "one target colour per ray"
        # per-ray scratch [N,13]: wsum | depth | image(3) | g_image(3) | g_ws | sqerr | (3 spare)
        pr = s["per_ray"]
        base = pr.data_ptr()
        wsum_p, depth_p, image_p, gimg_p, gws_p, sq_p = base, base + 4 * N, base + 8 * N, base + 20 * N, base + 32 * N, base + 36 * N
        # compositor + loss head in one launch: pred = image + (1 - ws) bg, squared error per ray, d(scaled loss)/d(image, ws)
        nb.call("ngp_composite_rays_train_forward_mse", sigma.data_ptr(), rgb.data_ptr(), deltas.data_ptr(), rays.data_ptr(), M, N,
                float(self.T_thresh), target.data_ptr(), float(self.bg), float(2.0 / (3.0 * self.R)), self.opt.scale_tensor().data_ptr(),
                wsum_p, depth_p, image_p, gimg_p, gws_p, sq_p)
        loss = pr.view(-1)[9 * N:10 * N].sum() * (1.0 / (3.0 * self.R))
        grads = s["grads"]
        grads.zero_()                                     # samples behind a ray's termination point receive no gradient
        g_rgb, g_sigma = grads.view(-1)[:3 * M].view(M, 3), grads.view(-1)[3 * M:]
        nb.call("ngp_composite_rays_train_backward", gws_p, gimg_p, sigma.data_ptr(),
                rgb.data_ptr(), deltas.data_ptr(), rays.data_ptr(), wsum_p, image_p, M, N, float(self.T_thresh),
                g_sigma.data_ptr(), g_rgb.data_ptr())
        if m.density_scale != 1:
            g_sigma = g_sigma * m.density_scale
        field_backward(stash["tensors"], stash["cfg"], stash["sinks"], g_sigma, g_rgb, side=side_or_none)
        if self._next is not None:                   # prefetch_point == "exchange": the march fills the communication bubble
            self.prefetch(*self._next)
            self._next = None
        self.opt.begin_exchange(self.group)          # async all-reduce of the fp16 sink (world size > 1)
        self.opt.finish_exchange()
        self.opt.apply()
        if self._pending:
            torch.cuda.current_stream().wait_stream(ms)
            self._pending = False
        self._ready[self._cur] = False
        self._cur = 1 - self._cur
        return loss

    @torch.no_grad()
    def __call__(self, rays_o, rays_d, target):
        """Unpipelined convenience form: march these rays now, then run the step on them."""
        self.march(rays_o, rays_d)
        return self.body(target)

    @torch.no_grad()
    def step_prefetched(self, target, next_rays_o, next_rays_d):
        """Steady-state form: the current slot was filled by the previous call (or by march()); consume it with `target` while
        the next step's rays are marched in the background."""
        self._next = (next_rays_o, next_rays_d)
        return self.body(target)


class EvalRenderer:
    """Full-frame inference (the eval branch of nerf/renderer.py run_cuda, :323-372) without a host round trip per iteration.

    Same kernels' arithmetic as the reference loop — march n_step samples per alive ray, evaluate the field, composite, drop finished
    rays, n_step = max(min(N / n_alive, 8), 1) — but the loop state (n_alive, n_step, M) lives in a device control block, alive rays
    are compacted on the device, the field is the fused inference pair (encoder -> sigma MLP -> exp ; SH + geo -> color MLP -> sigmoid,
    affine map, trunc_exp and sigmoid inside the kernels), and `block` iterations are captured in ONE CUDA graph that is replayed
    until the control block reports no alive rays (one 4-byte host read per block instead of a boolean-index sync per iteration)."""

    def __init__(self, model, n_rays, block=8, T_thresh=1e-4, dt_gamma=0.0, max_steps=1024, use_graph=True):
        self.m, self.N, self.block = model, int(n_rays), int(block)
        self.T_thresh, self.dt_gamma, self.max_steps, self.use_graph = T_thresh, dt_gamma, max_steps, use_graph
        dev = model.density_bitfield.device
        N = self.N
        Mmax = N + 128 + (128 - (N + 128) % 128)          # n_alive * n_step <= N, padded like the wrapper does
        f = dict(dtype=torch.float32, device=dev)
        self.ctrl = torch.zeros(8, dtype=torch.int32, device=dev)
        self.alive = [torch.empty(N, dtype=torch.int32, device=dev) for _ in range(2)]
        self.rays_o, self.rays_d = torch.empty(N, 3, **f), torch.empty(N, 3, **f)
        self.nears, self.fars, self.rays_t = torch.empty(N, **f), torch.empty(N, **f), torch.empty(N, **f)
        self.xyzs, self.dirs, self.deltas = torch.zeros(Mmax, 3, **f), torch.zeros(Mmax, 3, **f), torch.zeros(Mmax, 2, **f)
        self.h = torch.empty(Mmax, 16, dtype=torch.half, device=dev)
        self.sigma, self.rgb = torch.zeros(Mmax, **f), torch.zeros(Mmax, 3, **f)
        self.wsum, self.depth, self.image = torch.empty(N, **f), torch.empty(N, **f), torch.empty(N, 3, **f)
        self.Mmax = Mmax
        self.graph = None
        self.iterations = 0
        if model.density_scale != 1:
            raise RuntimeError("EvalRenderer: density_scale != 1 is not supported by the fused inference loop")

    def _iteration(self, cur):
        import _ngp_b200 as nb
        from nerf_fused import field_cfg
        from ngp_autograd import _half_table
        m, N = self.m, self.N
        enc = m.encoder
        a_in, a_out = self.alive[cur], self.alive[1 - cur]
        c = self.ctrl.data_ptr()
        nb.call("ngp_march_rays_dev", c, N, a_in.data_ptr(), self.rays_t.data_ptr(), self.rays_o.data_ptr(), self.rays_d.data_ptr(),
                float(m.bound), float(self.dt_gamma), int(self.max_steps), int(m.cascade), int(m.grid_size), m.density_bitfield.data_ptr(),
                self.fars.data_ptr(), self.xyzs.data_ptr(), self.dirs.data_ptr(), self.deltas.data_ptr(), None)
        nb.call("ngp_field_sigma_forward_dev", self.xyzs.data_ptr(), float(m.bound), c + 8, self._table.data_ptr(), enc.offsets.data_ptr(),
                enc.offsets.shape[0] - 1, float(np.log2(enc.per_level_scale)), int(enc.base_resolution), enc.gridtype_id,
                int(enc.align_corners), self._ws.data_ptr(), m.sigma_net.num_layers, self.Mmax, self.h.data_ptr(), self.sigma.data_ptr())
        nb.call("ngp_field_color_forward_dev", self.dirs.data_ptr(), self.h.data_ptr(), c + 8, self._wc.data_ptr(), m.color_net.num_layers,
                self.Mmax, self.rgb.data_ptr())
        nb.call("ngp_composite_rays_dev", c, N, float(self.T_thresh), a_in.data_ptr(), self.rays_t.data_ptr(), self.sigma.data_ptr(),
                self.rgb.data_ptr(), self.deltas.data_ptr(), self.wsum.data_ptr(), self.depth.data_ptr(), self.image.data_ptr())
        nb.call("ngp_compact_rays_dev", c, N, int(self.max_steps), a_in.data_ptr(), a_out.data_ptr())

    def _block(self):
        for k in range(self.block):
            self._iteration(k % 2)

    @torch.no_grad()
    def __call__(self, rays_o, rays_d, bg_color=1.0):
        import _ngp_b200 as nb
        from nerf_fused import field_cfg
        from ngp_autograd import _half_table, _half_param
        m, N = self.m, self.N
        assert self.block % 2 == 0, "an even block keeps the alive-list ping-pong aligned across replays"
        field_cfg(m.encoder, m.sigma_net, m.color_net, m.bound, False)       # validates the topology
        self.rays_o.copy_(rays_o.reshape(-1, 3)); self.rays_d.copy_(rays_d.reshape(-1, 3))
        # weights may have changed since the last frame: refresh the fp16 operands in place (graph-stable addresses)
        table = _half_table(m.encoder.embeddings)
        if getattr(self, "_table", None) is None or self._table.shape != table.shape:
            self._table = table.clone() if table is not getattr(m.encoder.embeddings, "_ngp_half_shadow", None) else table
            self._ws = _half_param(m.sigma_net.weights).clone()
            self._wc = _half_param(m.color_net.weights).clone()
        else:
            if self._table is not table:
                self._table.copy_(table)
            self._ws.copy_(_half_param(m.sigma_net.weights)); self._wc.copy_(_half_param(m.color_net.weights))
        nb.call("ngp_near_far_from_aabb", self.rays_o.data_ptr(), self.rays_d.data_ptr(), m.aabb_train.data_ptr(), N, float(m.min_near),
                self.nears.data_ptr(), self.fars.data_ptr())
        self.rays_t.copy_(self.nears)
        self.wsum.zero_(); self.depth.zero_(); self.image.zero_()
        nb.call("ngp_infer_init", N, self.alive[0].data_ptr(), self.ctrl.data_ptr())
        if self.use_graph and self.graph is None:
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._block()                       # warm-up outside capture (kernel attribute calls, allocator)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            # the warm-up consumed real iterations: restart the frame
            self.rays_t.copy_(self.nears); self.wsum.zero_(); self.depth.zero_(); self.image.zero_()
            nb.call("ngp_infer_init", N, self.alive[0].data_ptr(), self.ctrl.data_ptr())
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._block()
            # capture does not execute: the state is still the freshly initialised frame
        self.iterations = 0
        max_blocks = (self.max_steps + self.block - 1) // self.block + 1
        for _ in range(max_blocks):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._block()
            self.iterations += self.block
            if int(self.ctrl[0].item()) == 0:        # the only host read: once per `block` iterations
                break
        image = self.image + (1 - self.wsum).unsqueeze(-1) * bg_color
        depth = torch.clamp(self.depth - self.nears, min=0) / (self.fars - self.nears)
        return {"image": image, "depth": depth, "weights_sum": self.wsum.clone()}


def render_eval(model, rays_o, rays_d, bg_color=1.0, **kw):
    """One-shot convenience wrapper (keeps one EvalRenderer per model / ray count)."""
    n = rays_o.reshape(-1, 3).shape[0]
    cache = model.__dict__.setdefault("_eval_renderers", {})
    r = cache.get(n)
    if r is None:
        r = cache[n] = EvalRenderer(model, n, **kw)
    return r(rays_o, rays_d, bg_color)
