"""ngp_optim.py — fused mixed-precision optimizer step for the hot path's parameters.

One device pass per parameter tensor replaces the reference trainer's GradScaler.unscale_ + inf check +
torch.optim.Adam(betas=(0.9, 0.99), eps=1e-15) + gradient zeroing (nerf/utils.py:861-868, main_nerf.py:132) and the
per-forward fp32 -> fp16 cast of the hash table (gridencoder/grid.py:43-44).  Gradients are produced in fp16 directly into
one flat bucket (the `.grad` of the parameters is never materialised in fp32), which is also the single allreduce
payload under data parallelism (24.5 MB instead of 49 MB).  Loss-scale bookkeeping follows torch.amp.GradScaler and
lives on the device: there is no host synchronisation in `step()`.
"""
import torch
import torch.distributed as dist

import _ngp_b200 as _backend
from ngp_autograd import own_half_table, invalidate_half_table
from ngp_dp import segment_pieces


class FusedFieldOptimizer:
    """exchange="auto": at world size > 1 the gradient exchange runs over NVLink peer memory fused with a SHARDED optimizer pass
    (csrc/exchange.cu through ngp_dp.PeerExchange: reduce-scatter by peer loads, Adam on 1/world of the parameters, fp16 operand copies
    stored into every replica); "nccl": one NCCL all-reduce of the fp16 sink and the full optimizer pass on every rank (the baseline);
    with the peer path the fp32 masters / moments of the other ranks' shards are stale until gather_master()."""

    def __init__(self, encoder, sigma_net, color_net, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, init_scale=65536.0,
                 growth_factor=2.0, backoff_factor=0.5, growth_interval=2000, exchange="auto", group=None):
        self.encoder = encoder
        self.params = [encoder.embeddings, sigma_net.weights, color_net.weights]
        dev = self.params[0].device
        self.lr, self.betas, self.eps = lr, betas, eps
        self.growth, self.backoff, self.growth_interval = growth_factor, backoff_factor, growth_interval
        n = sum(p.numel() for p in self.params)
        self.group = group
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        assert exchange in ("auto", "peer", "nccl")
        self.px = None
        if world > 1 and exchange in ("auto", "peer") and dev.type == "cuda":
            try:
                from ngp_dp import PeerExchange
                self.px = PeerExchange(n, group=group, device=dev)
            except Exception as e:          # no peer access between the ranks' devices (or not one node): NCCL carries the exchange
                if exchange == "peer":
                    raise
                import warnings
                warnings.warn(f"FusedFieldOptimizer: peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL all-reduce")
                self.px = None
            if world > 1:
                # every rank must have taken the same decision
                ok = torch.tensor([1 if self.px is not None else 0], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                if int(ok.item()) == 0 and self.px is not None:
                    self.px.close()
                    self.px = None
        if self.px is not None:
            self.sink, self.shadow_flat = self.px.sink, self.px.shadow        # views of the peer-visible block
        else:
            self.sink = torch.zeros(n, dtype=torch.half, device=dev)          # flat fp16 gradient bucket
            self.shadow_flat = torch.zeros(n, dtype=torch.half, device=dev)   # flat fp16 operand copies (hash table, MLP weights)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.state = torch.zeros(8, dtype=torch.int32, device=dev)          # {scale, growth_tracker, found_inf, step, lr_scale, -, -, -}
        self.state[0:1].view(torch.float32).fill_(init_scale)
        self.state[4:5].view(torch.float32).fill_(1.0)
        self._exchange = None
        self._fused_args = None
        self.segments = []
        off = 0
        for p in self.params:
            k = p.numel()
            view = self.sink[off:off + k].view_as(p)
            p._ngp_grad_sink = view          # nerf_fused writes its fp16 gradients straight into this view
            # the fp16 kernel operand of every parameter is owned here: the Adam kernel rewrites it with every update
            own_half_table(p, into=self.shadow_flat[off:off + k].view_as(p))
            self.segments.append((p, off, k))
            off += k
        self.shadow = self.encoder.embeddings._ngp_half_shadow
        if self.px is not None:
            lo, hi = self.px.my_range
            self._pieces = segment_pieces([(o, k) for _, o, k in self.segments], lo, hi)
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)        # all shadows initialised before any rank's first exchange writes into them

    def scale_tensor(self):
        """Device scalar holding the current loss scale: `(loss * opt.scale_tensor()).backward()`."""
        return self.state[0:1].view(torch.float32)

    def set_lr_scale(self, factor):
        """LambdaLR-style schedule: the step size becomes lr * factor.  A device scalar, so it also takes effect in a captured graph."""
        self.state[4:5].view(torch.float32).fill_(float(factor))

    def refresh_shadow(self):
        """Call after anything else wrote the parameters (`.data` writes such as EMA copy_to()/restore(), load_state_dict)."""
        for p, off, k in self.segments:
            own_half_table(p, into=self.shadow_flat[off:off + k].view_as(p))

    def detach(self):
        for p in self.params:
            if hasattr(p, "_ngp_grad_sink"):
                del p._ngp_grad_sink
            invalidate_half_table(p)
        if self.px is not None:
            self.px.close()
            self.px = None

    @torch.no_grad()
    def _absorb_autograd_grads(self):
        """Gradients that arrived through autograd (module-by-module path, grad_total_variation) instead of the fp16 sink are folded
        into the sink, so nothing is silently dropped.  They carry the same loss scale (the caller scaled the loss)."""
        for p, off, k in self.segments:
            if p.grad is not None:
                self.sink[off:off + k].add_(p.grad.reshape(-1).to(torch.half))
                p.grad = None

    def begin_exchange(self, group=None):
        """NCCL path: launch the step's only exchange (sum-allreduce of the fp16 sink) without blocking the current stream: NCCL runs it
        on its own stream after the work queued so far; finish_exchange() makes the current stream wait for it.  Peer path: nothing to
        start — the exchange is part of apply()."""
        self._absorb_autograd_grads()
        group = group if group is not None else self.group
        if self.px is None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            self._exchange = dist.all_reduce(self.sink, op=dist.ReduceOp.SUM, group=group, async_op=True)

    def finish_exchange(self):
        if self._exchange is not None:
            self._exchange.wait()
            self._exchange = None

    @torch.no_grad()
    def step(self, group=None):
        if self._exchange is None:
            self.begin_exchange(group)
        self.finish_exchange()
        self.apply()

    @torch.no_grad()
    def apply(self):
        """inf check + Adam + fp16 shadow refresh + gradient zeroing + GradScaler update; on the peer path preceded by the reduction
        of this rank's shard and followed by the stores into every replica's shadow."""
        if self.px is not None:
            return self._apply_peer()
        _backend.call("ngp_optim_check_finite", self.sink.data_ptr(), 1, self.sink.numel(), self.state.data_ptr())
        for p, off, k in self.segments:
            _backend.call("ngp_optim_adam_step", p.data_ptr(), self.exp_avg.data_ptr() + 4 * off,
                          self.exp_avg_sq.data_ptr() + 4 * off, self.sink.data_ptr() + 2 * off, 1, self.shadow_flat.data_ptr() + 2 * off, k,
                          float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                          self.state.data_ptr(), 1)
        _backend.call("ngp_optim_scaler_update", self.state.data_ptr(), float(self.growth), float(self.backoff),
                      int(self.growth_interval))

    def _apply_peer(self):
        """Three launches (csrc/exchange.cu, fused form): [all buckets complete | reduce my shard] -> [non-finite flags OR-ed over the ranks |
        Adam on my shard, operand copies into every replica, bucket cleared | "my stores are done"] -> [all replicas complete | scaler]."""
        px = self.px
        lo, hi = px.my_range
        if self._fused_args is None:
            import ctypes
            n = len(self._pieces)
            P = (ctypes.c_void_p * 4)(*[self.segments[i][0].data_ptr() for i, _, _ in self._pieces] + [None] * (4 - n))
            U = ctypes.c_uint64 * 4
            seg = U(*[self.segments[i][1] for i, _, _ in self._pieces] + [0] * (4 - n))
            los = U(*[a for _, a, _ in self._pieces] + [0] * (4 - n))
            cnt = U(*[c for _, _, c in self._pieces] + [0] * (4 - n))
            self._fused_args = (P, seg, los, cnt, n)
        P, seg, los, cnt, n = self._fused_args
        tmo = 20000
        _backend.call("ngp_exchange_reduce_fused", px.pads, px.sinks, px.mc_sink, px.rank, px.world, lo, hi - lo, self.state.data_ptr(), tmo)
        _backend.call("ngp_exchange_adam_fused", px.pads, px.shadows, px.mc_shadow, px.rank, px.world, P, seg, los, cnt, n, self.exp_avg.data_ptr(),
                      self.exp_avg_sq.data_ptr(), self.sink.data_ptr(), lo, hi, self.sink.numel(), float(self.lr), float(self.betas[0]),
                      float(self.betas[1]), float(self.eps), self.state.data_ptr(), tmo)
        _backend.call("ngp_exchange_finish", px.pads, px.rank, px.world, self.state.data_ptr(), float(self.growth), float(self.backoff),
                      int(self.growth_interval), tmo)

    @torch.no_grad()
    def gather_master(self):
        """Peer path: every rank updates only its own shard of the fp32 masters and moments.  Collect all shards on all ranks (before a
        checkpoint, an evaluation through the fp32 parameters, or a switch of optimizer)."""
        if self.px is None:
            return
        px = self.px
        for r in range(px.world):
            lo, hi = px.bounds[r], px.bounds[r + 1]
            dist.broadcast(self.exp_avg[lo:hi], src=dist.get_global_rank(self.group, r) if self.group is not None else r, group=self.group)
            dist.broadcast(self.exp_avg_sq[lo:hi], src=dist.get_global_rank(self.group, r) if self.group is not None else r, group=self.group)
            for i, a, cnt in segment_pieces([(o, k) for _, o, k in self.segments], lo, hi):
                p, off, _ = self.segments[i]
                dist.broadcast(p.data.view(-1)[a - off:a - off + cnt], src=dist.get_global_rank(self.group, r) if self.group is not None else r,
                               group=self.group)

    # ---- checkpoint interchange with the reference trainer (nerf/utils.py:1015-1136) --------------------------------
    # The reference builds torch.optim.Adam(model.get_params(lr), betas=(0.9, 0.99), eps=1e-15) with the groups
    # [encoder], [sigma_net], [encoder_dir] (no parameters), [color_net] (network_ff.py:137-149) and a torch.amp.GradScaler.
    def state_dict(self):
        """torch.optim.Adam-format state dict for the reference's parameter-group layout (loadable by its trainer)."""
        self.gather_master()
        step = int(self.state[3].item())
        state = {}
        for i, (p, off, k) in enumerate(self.segments):
            state[i] = {"step": torch.tensor(float(step)),
                        "exp_avg": self.exp_avg[off:off + k].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(p).clone()}
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                     capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        groups = [dict(group, params=ids) for ids in ([0], [1], [], [2])]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        """Accepts the reference trainer's Adam state (parameters numbered in get_params order: embeddings, sigma, color)."""
        ids = [i for g in sd["param_groups"] for i in g["params"]]
        if len(ids) != len(self.segments):
            raise RuntimeError(f"FusedFieldOptimizer: checkpoint has {len(ids)} parameters, expected {len(self.segments)}")
        steps = set()
        for i, (p, off, k) in zip(ids, self.segments):
            st = sd["state"].get(i)
            if st is None:                       # parameter never stepped
                self.exp_avg[off:off + k].zero_(); self.exp_avg_sq[off:off + k].zero_(); steps.add(0)
                continue
            if st["exp_avg"].numel() != k:
                raise RuntimeError("FusedFieldOptimizer: optimizer state shape mismatch")
            self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) != 1:
            raise RuntimeError("FusedFieldOptimizer: parameters with different step counts are not supported")
        self.state[3] = steps.pop()
        g0 = sd["param_groups"][0]
        self.lr, self.betas, self.eps = g0["lr"], tuple(g0["betas"]), g0["eps"]

    def scaler_state_dict(self):
        """torch.amp.GradScaler.state_dict() format."""
        return {"scale": float(self.scale_tensor().item()), "growth_factor": self.growth, "backoff_factor": self.backoff,
                "growth_interval": self.growth_interval, "_growth_tracker": int(self.state[1].item())}

    def load_scaler_state_dict(self, sd):
        self.scale_tensor().fill_(float(sd["scale"]))
        self.state[1] = int(sd["_growth_tracker"])
        self.growth, self.backoff, self.growth_interval = sd["growth_factor"], sd["backoff_factor"], sd["growth_interval"]

