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


class FusedFieldOptimizer:
    def __init__(self, encoder, sigma_net, color_net, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, init_scale=65536.0,
                 growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.encoder = encoder
        self.params = [encoder.embeddings, sigma_net.weights, color_net.weights]
        dev = self.params[0].device
        self.lr, self.betas, self.eps = lr, betas, eps
        self.growth, self.backoff, self.growth_interval = growth_factor, backoff_factor, growth_interval
        n = sum(p.numel() for p in self.params)
        self.sink = torch.zeros(n, dtype=torch.half, device=dev)            # flat fp16 gradient bucket
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.state = torch.zeros(8, dtype=torch.int32, device=dev)          # {scale, growth_tracker, found_inf, step, lr_scale, -, -, -}
        self.state[0:1].view(torch.float32).fill_(init_scale)
        self.state[4:5].view(torch.float32).fill_(1.0)
        # the fp16 kernel operand of the hash table is owned here: the Adam kernel rewrites it with every update
        self.shadow = own_half_table(encoder.embeddings)
        self._exchange = None
        self.segments = []
        off = 0
        for p in self.params:
            k = p.numel()
            view = self.sink[off:off + k].view_as(p)
            p._ngp_grad_sink = view          # nerf_fused writes its fp16 gradients straight into this view
            self.segments.append((p, off, k))
            off += k

    def scale_tensor(self):
        """Device scalar holding the current loss scale: `(loss * opt.scale_tensor()).backward()`."""
        return self.state[0:1].view(torch.float32)

    def set_lr_scale(self, factor):
        """LambdaLR-style schedule: the step size becomes lr * factor.  A device scalar, so it also takes effect in a captured graph."""
        self.state[4:5].view(torch.float32).fill_(float(factor))

    def refresh_shadow(self):
        """Call after anything else wrote the hash table (`.data` writes such as EMA copy_to()/restore(), load_state_dict)."""
        self.shadow = own_half_table(self.encoder.embeddings)

    def detach(self):
        for p in self.params:
            if hasattr(p, "_ngp_grad_sink"):
                del p._ngp_grad_sink
        invalidate_half_table(self.encoder.embeddings)

    @torch.no_grad()
    def _absorb_autograd_grads(self):
        """Gradients that arrived through autograd (module-by-module path, grad_total_variation) instead of the fp16 sink are folded
        into the sink, so nothing is silently dropped.  They carry the same loss scale (the caller scaled the loss)."""
        for p, off, k in self.segments:
            if p.grad is not None:
                self.sink[off:off + k].add_(p.grad.reshape(-1).to(torch.half))
                p.grad = None

    def begin_exchange(self, group=None):
        """Launch the step's only exchange (sum-allreduce of the fp16 sink) without blocking the current stream: NCCL runs it on
        its own stream after the work queued so far; finish_exchange() makes the current stream wait for it."""
        self._absorb_autograd_grads()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
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
        """inf check + Adam + fp16 shadow refresh + gradient zeroing + GradScaler update on the (already reduced) sink."""
        _backend.call("ngp_optim_check_finite", self.sink.data_ptr(), 1, self.sink.numel(), self.state.data_ptr())
        for p, off, k in self.segments:
            shadow = self.shadow if p is self.encoder.embeddings else None    # refreshed in place by the kernel
            _backend.call("ngp_optim_adam_step", p.data_ptr(), self.exp_avg.data_ptr() + 4 * off,
                          self.exp_avg_sq.data_ptr() + 4 * off, self.sink.data_ptr() + 2 * off, 1, _backend.ptr(shadow), k,
                          float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                          self.state.data_ptr(), 1)
        _backend.call("ngp_optim_scaler_update", self.state.data_ptr(), float(self.growth), float(self.backoff),
                      int(self.growth_interval))

    # ---- checkpoint interchange with the reference trainer (nerf/utils.py:1015-1136) --------------------------------
    # The reference builds torch.optim.Adam(model.get_params(lr), betas=(0.9, 0.99), eps=1e-15) with the groups
    # [encoder], [sigma_net], [encoder_dir] (no parameters), [color_net] (network_ff.py:137-149) and a torch.amp.GradScaler.
    def state_dict(self):
        """torch.optim.Adam-format state dict for the reference's parameter-group layout (loadable by its trainer)."""
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

