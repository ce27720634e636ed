"""ngp_lazy.py — cross-module fusion behind the UNCHANGED module boundary ("lazy handle").

nerf/network_ff.py:51-74 (and sdf/netowrk_ff.py:37-46) call the drop-in modules one after the other:

    x = self.encoder(x, bound=self.bound)          # GridEncoder   -> [M,32] features
    h = self.sigma_net(x)                          # FFMLP         -> [M,16]
    d = self.encoder_dir(d)                        # SHEncoder     -> [M,16]
    h = torch.cat([d, geo_feat, p], dim=-1)        # torch         -> [M,32]
    h = self.color_net(h)                          # FFMLP         -> [M,3]

Run literally, the encoder features, the SH basis and the concatenated color input each make a round trip through HBM.  Instead,
GridEncoder / SHEncoder return a *deferred* tensor (a torch.Tensor wrapper subclass that carries the inputs and the producing
module): when the consumer is the drop-in FFMLP, it launches the fused kernel of csrc/ffmlp.cu (encoder -> tensor-core MLP from shared
memory: ngp_field_sigma_forward; SH + concat staging -> MLP: ngp_field_color_forward_ex) and the intermediate never exists; any
other consumer — a torch function, a method, an attribute that needs data — materialises the tensor with the ordinary op first, so
callers such as nerf/network.py (nn.Linear consumers) see exactly the eager result.  Deferred evaluation is used only where the fused
kernel is bit-identical in its encoder part to the eager op (3-D, 2 features/level, linear interpolation, fp16 table under autocast,
no gradient w.r.t. the coordinates); everything else stays eager.

Autograd: each fused call is one torch.autograd.Function over (coordinates, table, MLP weights) resp. (directions, geo features, MLP
weights); its backward runs the fused dgrad+wgrad kernel and, for the sigma net, the table scatter.  Gradients land in the parameters'
.grad (cast by autograd like the eager ops') or straight in the fp16 sinks of ngp_optim.FusedFieldOptimizer when that is installed.
"""
import numpy as np
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd
from torch.utils._pytree import tree_map

import _ngp_b200 as _backend
from ngp_autograd import _half_table, grid_encode, sh_encode

enabled = True            # module-level switch (tests / benchmarks turn deferred evaluation off to time the literal sequence)

_META = {"shape", "dtype", "device", "requires_grad", "is_cuda", "ndim", "layout", "is_leaf", "grad_fn", "names", "is_sparse",
         "is_quantized", "is_meta", "is_cpu", "is_nested", "is_mkldnn", "is_xpu", "is_mps", "is_vulkan", "is_ipu", "is_xla", "is_mtia",
         "is_maia", "is_sparse_csr", "output_nr", "_version", "_base", "grad", "retains_grad"}
_META_METHODS = {"size", "dim", "numel", "nelement", "ndimension", "is_floating_point", "is_complex", "get_device", "element_size",
                 "is_contiguous", "stride", "storage_offset", "__len__"}


def _is_meta_call(func):
    name = getattr(func, "__name__", "")
    if name == "__get__":
        return getattr(getattr(func, "__self__", None), "__name__", None) in _META
    return name in _META_METHODS


class Deferred(torch.Tensor):
    """Base of the deferred tensors: metadata (shape / dtype / device) is real, data is produced on first use."""

    @staticmethod
    def _wrap(cls, shape, dtype, device):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(shape), dtype=dtype, device=device, requires_grad=False)
        t._grad_mode = torch.is_grad_enabled()       # the producing call's grad mode travels with the tensor, like its autocast state
        return t

    def materialize(self):
        """The eager op, run under the autocast state AND the grad mode the producing module saw (the consumer may sit outside the
        `autocast` / `no_grad` region: `with torch.no_grad(): f = encoder(x)` followed by `f.float()` must not build a graph)."""
        if self._value is None:
            with torch.autocast('cuda', dtype=torch.half, enabled=True), torch.set_grad_enabled(getattr(self, "_grad_mode", True)):
                self._value = self._compute()
        return self._value

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if _is_meta_call(func):
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        special = cls._intercept(func, args, kwargs)
        if special is not NotImplemented:
            return special
        args, kwargs = tree_map(lambda a: a.materialize() if isinstance(a, Deferred) else a, (args, kwargs))
        return func(*args, **kwargs)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # safety net: an ATen op reached the dispatcher with a deferred tensor (a code path that skipped __torch_function__)
        args, kwargs = tree_map(lambda a: a.materialize() if isinstance(a, Deferred) else a, (args, kwargs or {}))
        return func(*args, **kwargs)

    @classmethod
    def _intercept(cls, func, args, kwargs):
        return NotImplemented

    def __repr__(self):
        return f"{type(self).__name__}(shape={tuple(self.shape)}, dtype={self.dtype}, device={self.device}, pending={self._value is None})"


class DeferredGridFeatures(Deferred):
    """GridEncoder(x, bound) not yet evaluated: `coords` are the world coordinates [..., 3]."""

    @staticmethod
    def make(encoder, coords, bound):
        lead = list(coords.shape[:-1])
        t = Deferred._wrap(DeferredGridFeatures, lead + [encoder.output_dim], torch.half, coords.device)
        t._encoder, t._coords, t._bound, t._value = encoder, coords, bound, None
        return t

    def _compute(self):
        return self._encoder._forward_eager(self._coords, self._bound)


class DeferredSH(Deferred):
    """SHEncoder(d, size) not yet evaluated."""

    @staticmethod
    def make(encoder, dirs, size):
        lead = list(dirs.shape[:-1])
        t = Deferred._wrap(DeferredSH, lead + [encoder.output_dim], torch.float32, dirs.device)
        t._encoder, t._dirs, t._size, t._value = encoder, dirs, size, None
        return t

    def _compute(self):
        return self._encoder._forward_eager(self._dirs, self._size)

    @classmethod
    def _intercept(cls, func, args, kwargs):
        # torch.cat([SH(d), geo_feat, p], dim=-1)  (network_ff.py:68): keep the three parts apart for the color FFMLP
        if func is torch.cat and args and isinstance(args[0], (list, tuple)) and len(args[0]) == 3:
            parts = args[0]
            dim = kwargs.get("dim", args[1] if len(args) > 1 else 0)
            sh, geo, pad = parts
            if (isinstance(sh, DeferredSH) and sh._value is None and not isinstance(geo, Deferred) and not isinstance(pad, Deferred)
                    and torch.is_tensor(geo) and torch.is_tensor(pad) and sh.dim() == 2 and geo.dim() == 2 and pad.dim() == 2
                    and dim in (-1, 1) and sh.shape[1] == 16 and geo.shape[1] == 15 and pad.shape[1] == 1
                    and geo.shape[0] == sh.shape[0] and pad.shape[0] == sh.shape[0] and sh._size == 1
                    and geo.dtype == torch.half and geo.is_cuda and not pad.requires_grad):
                return DeferredColorInput.make(sh, geo, pad)
        return NotImplemented


class DeferredColorInput(Deferred):
    """cat([SH4(d), geo_feat [M,15], pad [M,1]], -1) not yet evaluated."""

    @staticmethod
    def make(sh, geo, pad):
        # the SH part is float32: cat promotes (also under autocast, where cat is a "promote to widest" op)
        t = Deferred._wrap(DeferredColorInput, [sh.shape[0], 32], torch.float32, geo.device)
        t._sh, t._geo, t._pad, t._value = sh, geo, pad, None
        return t

    def _compute(self):
        return torch.cat([self._sh.materialize(), self._geo, self._pad], dim=-1)


# ------------------------------------------------------------------------------------------------ fused ops
def _grid_ok(enc, coords):
    return (enabled and coords.is_cuda and enc.input_dim == 3 and enc.level_dim == 2 and enc.interp_id == 0
            and enc.num_levels % 4 == 0 and enc.num_levels <= 32 and torch.is_autocast_enabled('cuda')
            and torch.get_autocast_dtype('cuda') == torch.half
            and not (coords.requires_grad and torch.is_grad_enabled()) and coords.shape[-1] == 3 and coords.numel() > 0)


def defer_grid(enc, coords, bound):
    """GridEncoder.forward's hook: a deferred tensor when the fused consumer would reproduce the eager op exactly, else None."""
    return DeferredGridFeatures.make(enc, coords, bound) if _grid_ok(enc, coords) else None


def defer_sh(enc, dirs, size):
    ok = (enabled and dirs.is_cuda and enc.degree == 4 and dirs.shape[-1] == 3 and dirs.dim() == 2 and size == 1
          and not (dirs.requires_grad and torch.is_grad_enabled()) and dirs.numel() > 0 and torch.is_autocast_enabled('cuda')
          and torch.get_autocast_dtype('cuda') == torch.half)
    return DeferredSH.make(enc, dirs, size) if ok else None


class _GridMLPFn(Function):
    """h = FFMLP(GridEncoder(x)) in one kernel (ngp_field_sigma_forward); backward = fused MLP backward + table scatter."""

    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, coords, embeddings, offsets, weights, cfg):
        bound, pls, H, gridtype, align, nl, out_dim, inference = cfg
        x01 = ((coords.float() + bound) / (2 * bound)).contiguous().view(-1, 3)      # GridEncoder.forward's affine map (grid.py:149)
        table = _half_table(embeddings)
        w = _half_table(weights).contiguous()          # owner-maintained fp16 copy when the fused optimizer holds one, else a cast
        M, L, S, dev = x01.shape[0], offsets.shape[0] - 1, float(np.log2(pls)), x01.device
        h = torch.empty(M, 16, dtype=torch.half, device=dev)
        feat = fb = None
        if not inference:
            feat = torch.empty(M, 2 * L, dtype=torch.half, device=dev)
            fb = torch.empty(nl, M, 64, dtype=torch.half, device=dev)
        _backend.call("ngp_field_sigma_forward", x01.data_ptr(), table.data_ptr(), offsets.data_ptr(), L, S, int(H), gridtype, int(align),
                      w.data_ptr(), nl, M, int(not inference), _backend.ptr(feat), _backend.ptr(fb), h.data_ptr(), None)
        if not inference:
            ctx.save_for_backward(x01, offsets, w, feat, fb)
            ctx.cfg = (L, S, int(H), gridtype, int(align), nl, M, tuple(table.shape))
            ctx.sinks = (getattr(embeddings, "_ngp_grad_sink", None), getattr(weights, "_ngp_grad_sink", None))
        return h if out_dim == 16 else h[:, :out_dim]

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        x01, offsets, w, feat, fb = ctx.saved_tensors
        L, S, H, gridtype, align, nl, M, table_shape = ctx.cfg
        sink_t, sink_w = ctx.sinks
        dev = x01.device
        g = grad.half()
        if g.shape[1] != 16:                                 # padded output columns carry no gradient
            g = torch.nn.functional.pad(g, (0, 16 - g.shape[1]))
        g = g.contiguous()
        d_feat = torch.empty(M, 2 * L, dtype=torch.half, device=dev)
        gw = sink_w if sink_w is not None else torch.empty_like(w)
        nb = _backend.load().ngp_ffmlp_backward_workspace_bytes(M, 2 * L, 16, 64, nl)
        wk = torch.empty(nb // 4, dtype=torch.float32, device=dev)
        _backend.call("ngp_ffmlp_backward", g.data_ptr(), feat.data_ptr(), w.data_ptr(), fb.data_ptr(), M, 2 * L, 16, 64, nl, 0, 6, 1, None,
                      d_feat.data_ptr(), gw.data_ptr(), wk.data_ptr(), nb)
        g_table = sink_t if sink_t is not None else torch.zeros(table_shape, dtype=torch.half, device=dev)
        _backend.call("ngp_grid_encode_backward", d_feat.data_ptr(), x01.data_ptr(), None, offsets.data_ptr(), g_table.data_ptr(), M, 3, 2,
                      L, S, H, None, None, gridtype, align, 0, 1, 0)
        return None, (None if sink_t is not None else g_table), None, (None if sink_w is not None else gw), None


def grid_mlp(lazy, mlp):
    """FFMLP.forward's hook for a DeferredGridFeatures input; returns None when the fused kernel does not cover the configuration."""
    enc = lazy._encoder
    if not (enabled and lazy._value is None and mlp.hidden_dim == 64 and mlp.activation == 0 and mlp.output_activation == 6
            and mlp.input_dim == enc.output_dim and mlp.padded_output_dim == 16 and 2 <= mlp.num_layers <= 5
            and getattr(lazy, "_grad_mode", True) == torch.is_grad_enabled()):       # producer under no_grad, consumer not: eager semantics
        return None
    inference = not (mlp.training and torch.is_grad_enabled())
    cfg = (float(lazy._bound), float(enc.per_level_scale), int(enc.base_resolution), enc.gridtype_id, bool(enc.align_corners),
           mlp.num_layers, mlp.output_dim, inference)
    lead = list(lazy.shape[:-1])
    out = _GridMLPFn.apply(lazy._coords, enc.embeddings, enc.offsets, mlp.weights, cfg)
    return out.view(lead + [mlp.output_dim])


def _rows_of_16(geo):
    """geo [M,15] half that is columns 1..15 of a row-major [M,16] half buffer (h[..., 1:] of the sigma net's output): returns the
    address of that buffer's first element, else None."""
    if geo.dtype != torch.half or geo.dim() != 2 or geo.shape[1] != 15 or geo.stride() != (16, 1):
        return None
    ptr = geo.data_ptr() - 2
    return ptr if ptr % 32 == 0 else None


class _ColorMLPFn(Function):
    """y = FFMLP(cat([SH4(d), geo, pad])) in one kernel (ngp_field_color_forward_ex); geo is read in place from the sigma net's
    [M,16] output; backward = fused dgrad+wgrad (ngp_field_color_backward_ex) -> dL/d geo."""

    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, dirs, geo, pad, weights, cfg):
        nl, out_dim, inference = cfg
        d = dirs.float().contiguous()
        p = pad.half().contiguous().view(-1)
        w = _half_table(weights).contiguous()          # owner-maintained fp16 copy when the fused optimizer holds one, else a cast
        M, dev = d.shape[0], d.device
        h_ptr = _rows_of_16(geo)
        y = torch.empty(M, 16, dtype=torch.half, device=dev)
        fb = None if inference else torch.empty(nl, M, 64, dtype=torch.half, device=dev)
        _backend.call("ngp_field_color_forward_ex", d.data_ptr(), h_ptr, p.data_ptr(), w.data_ptr(), nl, M, int(not inference),
                      _backend.ptr(fb), None, y.data_ptr())
        if not inference:
            ctx.save_for_backward(d, geo, p, w, fb)
            ctx.cfg = (nl, M, geo.requires_grad)
            ctx.sink = getattr(weights, "_ngp_grad_sink", None)
        return y[:, :out_dim]

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        d, geo, p, w, fb = ctx.saved_tensors
        nl, M, want_geo = ctx.cfg
        sink = ctx.sink
        dev = d.device
        g = grad.half()
        if g.shape[1] != 3 or not g.is_contiguous():
            g3 = torch.zeros(M, 3, dtype=torch.half, device=dev)
            g3[:, :min(3, g.shape[1])] = g[:, :3]
            g = g3
        dys = torch.empty(M, 16, dtype=torch.half, device=dev)
        gw = sink if sink is not None else torch.empty_like(w)
        nb = _backend.load().ngp_ffmlp_backward_workspace_bytes(M, 32, 16, 64, nl)
        wk = torch.empty(nb // 4, dtype=torch.float32, device=dev)
        _backend.call("ngp_field_color_backward_ex", None, None, g.data_ptr(), None, _rows_of_16(geo), d.data_ptr(), p.data_ptr(),
                      w.data_ptr(), fb.data_ptr(), nl, M, dys.data_ptr(), gw.data_ptr(), wk.data_ptr(), nb, 0)
        return None, (dys[:, 1:] if want_geo else None), None, (None if sink is not None else gw), None


def color_mlp(lazy, mlp):
    """FFMLP.forward's hook for a DeferredColorInput; None when the fused kernel does not cover the configuration."""
    geo = lazy._geo
    if not (enabled and lazy._value is None and mlp.hidden_dim == 64 and mlp.activation == 0 and mlp.output_activation == 6
            and mlp.input_dim == 32 and mlp.padded_output_dim == 16 and mlp.output_dim == 3 and 2 <= mlp.num_layers <= 5
            and _rows_of_16(geo) is not None and getattr(lazy, "_grad_mode", True) == torch.is_grad_enabled()):
        return None
    inference = not (mlp.training and torch.is_grad_enabled())
    return _ColorMLPFn.apply(lazy._sh._dirs, geo, lazy._pad, mlp.weights, (mlp.num_layers, mlp.output_dim, inference))
