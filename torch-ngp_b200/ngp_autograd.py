"""ngp_autograd.py — the differentiable ops of the hot path as torch.autograd Functions over the C ABI (include/ngp_b200.h).

One place for the four encoder / MLP ops the drop-in packages re-export under the reference's names:

    grid_encode     gridencoder/grid.py:27-93 of the reference          -> ngp_grid_encode_forward / _backward
    ffmlp_forward   ffmlp/ffmlp.py:15-86                                -> ngp_ffmlp_forward / _inference / _backward
    sh_encode       shencoder/sphere_harmonics.py:14-57                 -> ngp_sh_encode_forward / _backward
    freq_encode     freqencoder/freq.py:15-53                           -> ngp_freq_encode_forward / _backward

Call signatures (positional order, defaults) and autocast contracts are the reference's: the grid op manages autocast itself
(half table, float coordinates), the MLP casts its inputs to half, SH / frequency encodings are forced to float32.  Differences
that are invisible to callers: kernels run on the current stream, the fp16 copy of the hash table is kept current by the fused
optimizer when that owns the parameters (re-cast per forward otherwise, as in the reference),
the MLP backward needs no [layers, B, hidden] scratch unless the net is deeper than the fused kernel supports.  No CPU path.
"""
from collections import namedtuple

import numpy as np
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

import _ngp_b200 as _backend

_call, _ptr = _backend.call, _backend.ptr


def _half_table(embeddings):
    """fp16 copy of the table for the kernels (reference gridencoder/grid.py:43-44 casts on every forward).

    A cached copy is reused ONLY while ngp_optim.FusedFieldOptimizer owns it: that optimizer's Adam kernel rewrites the shadow in
    place with every parameter update, so it is always current for updates made through it.  Any other writer (torch optimizers,
    `.data` writes such as torch_ema copy_to()/restore(), reset_parameters, load_state_dict) is invisible to a cache — `.data`
    writes do not even bump the autograd version counter — so without an owner the table is re-cast on every call, as the reference
    does.  Owners that let someone else write the parameter call invalidate_half_table() (FusedFieldOptimizer.refresh_shadow)."""
    hit = getattr(embeddings, "_ngp_half_shadow", None)
    if hit is not None and hit.shape == embeddings.shape and hit.device == embeddings.device:
        return hit
    return embeddings.detach().to(torch.half)


def own_half_table(embeddings, into=None):
    """Install (or refresh) the owner-maintained fp16 shadow of `embeddings`; returns it.  `into` (optional): a half tensor of the
    parameter's shape that becomes the shadow (the fused optimizer keeps all shadows in one flat, peer-visible buffer)."""
    if into is not None:
        assert into.shape == embeddings.shape and into.dtype == torch.half and into.device == embeddings.device
        with torch.no_grad():
            into.copy_(embeddings.detach())
        embeddings._ngp_half_shadow = into
        return into
    hit = getattr(embeddings, "_ngp_half_shadow", None)
    if hit is not None and hit.shape == embeddings.shape and hit.device == embeddings.device:
        hit.copy_(embeddings.detach())
        return hit
    half = embeddings.detach().to(torch.half)
    embeddings._ngp_half_shadow = half
    return half


def _half_param(p):
    """fp16 kernel operand of any hot-path parameter (hash table or flat FFMLP weight vector): the owner-maintained shadow when one
    exists (always current, see _half_table), otherwise a fresh cast as the reference does (ffmlp.py:18 custom_fwd cast)."""
    return _half_table(p).contiguous()


def invalidate_half_table(embeddings):
    if hasattr(embeddings, "_ngp_half_shadow"):
        del embeddings._ngp_half_shadow


def sync_half_table(embeddings):
    """After an out-of-band write to `embeddings` (reset_parameters, load_state_dict): refresh an owner-maintained shadow in place
    (its address stays valid for captured graphs); drop it if it no longer matches the parameter's shape / device."""
    hit = getattr(embeddings, "_ngp_half_shadow", None)
    if hit is None:
        return
    if hit.shape == embeddings.shape and hit.device == embeddings.device:
        with torch.no_grad():
            hit.copy_(embeddings.detach())
    else:
        invalidate_half_table(embeddings)


# ------------------------------------------------------------------------------------------------ hash grid
_GridCfg = namedtuple("_GridCfg", "B D C L S H gridtype align interp dtype table_shape table_dtype")


class GridEncodeFn(Function):
    """x [B,D] in [0,1] (float32), table [entries, C], offsets [L+1] int32 -> features [B, L*C] in the table's dtype."""

    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0):
        _backend.require_cuda(inputs, embeddings, offsets)
        if inputs.dtype != torch.float32:
            raise RuntimeError("grid_encode: inputs must be float32 (reference: inputs.data_ptr<float>())")
        x = inputs.contiguous()
        # under autocast the kernels read a half-precision table (reference grid.py:41-44); coordinates stay float
        use_half = torch.is_autocast_enabled('cuda') and embeddings.shape[1] % 2 == 0
        table = _half_table(embeddings) if use_half else embeddings.detach().contiguous()
        if table.dtype not in (torch.float32, torch.float16):
            raise RuntimeError("grid_encode: embeddings must be float32 or float16")
        cfg = _GridCfg(B=x.shape[0], D=x.shape[1], C=table.shape[1], L=offsets.shape[0] - 1, S=float(np.log2(per_level_scale)),
                       H=int(base_resolution), gridtype=gridtype, align=int(align_corners), interp=interpolation,
                       dtype=1 if table.dtype == torch.float16 else 0, table_shape=table.shape, table_dtype=table.dtype)
        feats = torch.empty(cfg.B, cfg.L * cfg.C, device=x.device, dtype=table.dtype)
        jac = torch.empty(cfg.B, cfg.L * cfg.D * cfg.C, device=x.device, dtype=table.dtype) if calc_grad_inputs else None
        _call("ngp_grid_encode_forward", x.data_ptr(), table.data_ptr(), offsets.data_ptr(), feats.data_ptr(), cfg.B, cfg.D, cfg.C,
              cfg.L, cfg.S, cfg.H, _ptr(jac), cfg.gridtype, cfg.align, cfg.interp, cfg.dtype, 0)
        ctx.save_for_backward(x, offsets, jac)
        ctx.cfg = cfg
        return feats

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        x, offsets, jac = ctx.saved_tensors
        cfg = ctx.cfg
        g = grad.contiguous().to(cfg.table_dtype)
        d_table = torch.zeros(cfg.table_shape, device=g.device, dtype=cfg.table_dtype)      # scatter-add target
        d_x = torch.zeros_like(x, dtype=cfg.table_dtype) if jac is not None else None
        _call("ngp_grid_encode_backward", g.data_ptr(), x.data_ptr(), None, offsets.data_ptr(), d_table.data_ptr(), cfg.B, cfg.D,
              cfg.C, cfg.L, cfg.S, cfg.H, _ptr(jac), _ptr(d_x), cfg.gridtype, cfg.align, cfg.interp, cfg.dtype, 0)
        return (None if d_x is None else d_x.to(x.dtype)), d_table, None, None, None, None, None, None, None


grid_encode = GridEncodeFn.apply


# ------------------------------------------------------------------------------------------------ fully fused MLP
_MlpCfg = namedtuple("_MlpCfg", "B n_in n_out width layers act out_act want_dx")
FUSED_BACKWARD_MAX_MATMULS = 6      # deeper nets take the two-kernel path, which needs the dL/d(pre-activation) scratch in HBM


class FFMLPFn(Function):
    """y = MLP(x): x [B, n_in] half, weights = flat half vector of [out,in] matrices (ffmlp.cu:631-634) -> [B, n_out] half."""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.half)
    def forward(ctx, inputs, weights, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation,
                inference=False, calc_grad_inputs=False):
        _backend.require_cuda(inputs, weights)
        if inputs.dtype != torch.half or weights.dtype != torch.half:
            # outside autocast the reference's CHECK_IS_HALF raises; say what to do about it
            raise RuntimeError("ffmlp: inputs and weights must be half (run under torch.autocast or cast explicitly)")
        x, w = inputs.contiguous(), weights.contiguous()
        cfg = _MlpCfg(B=x.shape[0], n_in=input_dim, n_out=output_dim, width=hidden_dim, layers=num_layers, act=activation,
                      out_act=output_activation, want_dx=bool(calc_grad_inputs))
        y = torch.empty(cfg.B, cfg.n_out, device=x.device, dtype=x.dtype)
        if inference:
            _call("ngp_ffmlp_inference", x.data_ptr(), w.data_ptr(), cfg.B, cfg.n_in, cfg.n_out, cfg.width, cfg.layers, cfg.act,
                  cfg.out_act, None, y.data_ptr())
            return y
        stash = torch.empty(cfg.layers, cfg.B, cfg.width, device=x.device, dtype=x.dtype)     # post-activation hidden states
        _call("ngp_ffmlp_forward", x.data_ptr(), w.data_ptr(), cfg.B, cfg.n_in, cfg.n_out, cfg.width, cfg.layers, cfg.act, cfg.out_act,
              stash.data_ptr(), y.data_ptr())
        ctx.save_for_backward(x, w, stash)
        ctx.cfg = cfg
        return y

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        x, w, stash = ctx.saved_tensors
        cfg = ctx.cfg
        g = grad.contiguous().half()
        d_x = torch.empty_like(x) if cfg.want_dx else None
        d_w = torch.empty_like(w)
        scratch = (torch.empty(cfg.layers, cfg.B, cfg.width, device=g.device, dtype=g.dtype)
                   if cfg.layers + 1 > FUSED_BACKWARD_MAX_MATMULS else None)
        nbytes = _backend.load().ngp_ffmlp_backward_workspace_bytes(cfg.B, cfg.n_in, cfg.n_out, cfg.width, cfg.layers)
        workspace = torch.empty(nbytes // 4, device=g.device, dtype=torch.float32)
        _call("ngp_ffmlp_backward", g.data_ptr(), x.data_ptr(), w.data_ptr(), stash.data_ptr(), cfg.B, cfg.n_in, cfg.n_out, cfg.width,
              cfg.layers, cfg.act, cfg.out_act, int(cfg.want_dx), _ptr(scratch), _ptr(d_x), d_w.data_ptr(), workspace.data_ptr(), nbytes)
        return (d_x, d_w) + (None,) * 8


ffmlp_forward = FFMLPFn.apply


# ------------------------------------------------------------------------------------------------ spherical harmonics
class SHEncodeFn(Function):
    """unit directions [B,3] -> real SH basis [B, degree^2] (float32; degree 1..8)."""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        _backend.require_cuda(inputs)
        d = inputs.contiguous().float()
        B, n_in = d.shape
        basis = torch.empty(B, degree ** 2, dtype=d.dtype, device=d.device)
        jac = torch.empty(B, n_in * degree ** 2, dtype=d.dtype, device=d.device) if calc_grad_inputs else None
        _call("ngp_sh_encode_forward", d.data_ptr(), basis.data_ptr(), B, n_in, degree, _ptr(jac))
        ctx.save_for_backward(d, jac)
        ctx.shape = (B, n_in, degree)
        return basis

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        d, jac = ctx.saved_tensors
        if jac is None:
            return None, None, None
        B, n_in, degree = ctx.shape
        g = grad.contiguous().float()
        d_dir = torch.zeros_like(d)
        _call("ngp_sh_encode_backward", g.data_ptr(), d.data_ptr(), B, n_in, degree, jac.data_ptr(), d_dir.data_ptr())
        return d_dir, None, None


sh_encode = SHEncodeFn.apply


# ------------------------------------------------------------------------------------------------ frequency encoding
class FreqEncodeFn(Function):
    """x [B,D] -> [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] [B, D + 2*degree*D] (float32)."""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, output_dim):
        _backend.require_cuda(inputs)
        x = inputs.contiguous().float()
        B, n_in = x.shape
        enc = torch.empty(B, output_dim, dtype=x.dtype, device=x.device)
        _call("ngp_freq_encode_forward", x.data_ptr(), B, n_in, degree, output_dim, enc.data_ptr())
        ctx.save_for_backward(enc)
        ctx.shape = (B, n_in, degree, output_dim)
        return enc

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        (enc,) = ctx.saved_tensors
        B, n_in, degree, output_dim = ctx.shape
        g = grad.contiguous().float()
        d_x = torch.empty(B, n_in, dtype=enc.dtype, device=enc.device)
        _call("ngp_freq_encode_backward", g.data_ptr(), enc.data_ptr(), B, n_in, degree, output_dim, d_x.data_ptr())
        return d_x, None, None


freq_encode = FreqEncodeFn.apply
