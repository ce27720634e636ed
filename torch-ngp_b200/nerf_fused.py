"""nerf_fused.py — optional fused evaluation of the NeRF field (hash grid -> sigma MLP -> trunc_exp ; SH (+) geo ->
color MLP -> sigmoid) on top of the fused C-ABI extensions (include/ngp_b200.h, ngp_field_*).

Same mathematics, parameters and rounding points as the module-by-module path of nerf/network_ff.py:51-74 (reference)
— encoder features bit-identical, MLP outputs identical (same kernels), sigma/rgb within 1 ulp of torch.exp /
torch.sigmoid — but the encoder output feeds the tensor-core MLP from shared memory, SH + concat + casts happen in
the color kernel's input staging, and the backward runs sigmoid/cat/trunc_exp gradients inside the MLP backward.
A caller opts in with one line:   sigma, rgb = fused_field(self.encoder, self.sigma_net, self.color_net, x, d, self.bound)
"""
import numpy as np
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

import _ngp_b200 as _backend
from gridencoder.grid import _half_table


def field_forward(xyzs, dirs, embeddings, offsets, sigma_w, color_w, cfg):
    """Raw (autograd-free) fused field evaluation.  Returns sigma [M] f32, rgb [M,3] f32 and the backward stash (or None)."""
    bound, pls, H, gridtype, align_corners, nl_s, nl_c, training = cfg
    _backend.require_cuda(xyzs, dirs, embeddings, sigma_w, color_w)
    x01 = ((xyzs.float() + bound) / (2 * bound)).contiguous()       # GridEncoder.forward's affine map (grid.py:149)
    dirs = dirs.float().contiguous()
    table = _half_table(embeddings)
    ws = sigma_w.detach().half().contiguous()
    wc = color_w.detach().half().contiguous()
    M = x01.shape[0]
    L = offsets.shape[0] - 1
    S = float(np.log2(pls))
    dev = x01.device
    h = torch.empty(M, 16, dtype=torch.half, device=dev)
    sigma = torch.empty(M, dtype=torch.float32, device=dev)
    rgb = torch.empty(M, 3, dtype=torch.float32, device=dev)
    feat = fb_s = fb_c = None
    if training:
        feat = torch.empty(M, 2 * L, dtype=torch.half, device=dev)
        fb_s = torch.empty(nl_s, M, 64, dtype=torch.half, device=dev)
        fb_c = torch.empty(nl_c, M, 64, dtype=torch.half, device=dev)
    _backend.call("ngp_field_sigma_forward", x01.data_ptr(), table.data_ptr(), offsets.data_ptr(), L, S, int(H),
                  gridtype, int(align_corners), ws.data_ptr(), nl_s, M, int(training), _backend.ptr(feat),
                  _backend.ptr(fb_s), h.data_ptr(), sigma.data_ptr())
    _backend.call("ngp_field_color_forward", dirs.data_ptr(), h.data_ptr(), wc.data_ptr(), nl_c, M, int(training),
                  _backend.ptr(fb_c), rgb.data_ptr())
    stash = None
    if training:
        stash = dict(tensors=(x01, dirs, offsets, ws, wc, feat, fb_s, fb_c, h, rgb),
                     cfg=(L, S, int(H), gridtype, int(align_corners), nl_s, nl_c, M, tuple(table.shape)),
                     # optional fp16 gradient sinks installed by ngp_optim.FusedFieldOptimizer (bypass fp32 .grad accumulation)
                     sinks=tuple(getattr(p, "_ngp_grad_sink", None) for p in (embeddings, sigma_w, color_w)))
    return sigma, rgb, stash


def field_backward(tensors, cfg, sinks, d_sigma, d_rgb):
    """Raw backward of field_forward: returns (g_table, gw_sigma, gw_color) — the sinks themselves when installed."""
    x01, dirs, offsets, ws, wc, feat, fb_s, fb_c, h, rgb = tensors
    L, S, H, gridtype, align_corners, nl_s, nl_c, M, table_shape = cfg
    dev = x01.device
    d_sigma = d_sigma.float().contiguous()
    d_rgb = d_rgb.float().contiguous()
    lib = _backend.load()
    sink_t, sink_s, sink_c = sinks
    # color net (+ sigmoid, cat, trunc_exp gradients) -> dL/d(sigma-net output)
    dys = torch.empty(M, 16, dtype=torch.half, device=dev)
    gw_c = sink_c if sink_c is not None else torch.empty_like(wc)
    nb_c = lib.ngp_ffmlp_backward_workspace_bytes(M, 32, 16, 64, nl_c)
    wk_c = torch.empty(nb_c // 4, dtype=torch.float32, device=dev)
    _backend.call("ngp_field_color_backward", d_rgb.data_ptr(), rgb.data_ptr(), d_sigma.data_ptr(), h.data_ptr(),
                  dirs.data_ptr(), wc.data_ptr(), fb_c.data_ptr(), nl_c, M, dys.data_ptr(), gw_c.data_ptr(),
                  wk_c.data_ptr(), nb_c)
    # sigma net -> dL/d(features)
    d_feat = torch.empty(M, 2 * L, dtype=torch.half, device=dev)
    gw_s = sink_s if sink_s is not None else torch.empty_like(ws)
    nb_s = lib.ngp_ffmlp_backward_workspace_bytes(M, 2 * L, 16, 64, nl_s)
    wk_s = torch.empty(nb_s // 4, dtype=torch.float32, device=dev)
    _backend.call("ngp_ffmlp_backward", dys.data_ptr(), feat.data_ptr(), ws.data_ptr(), fb_s.data_ptr(), M, 2 * L, 16, 64,
                  nl_s, 0, 6, 1, None, d_feat.data_ptr(), gw_s.data_ptr(), wk_s.data_ptr(), nb_s)
    # hash-table scatter-add (a sink is kept zeroed by the optimizer kernel; otherwise a fresh zero table as in grid.py:77)
    g_table = sink_t if sink_t is not None else torch.zeros(table_shape, dtype=torch.half, device=dev)
    _backend.call("ngp_grid_encode_backward", d_feat.data_ptr(), x01.data_ptr(), None, offsets.data_ptr(),
                  g_table.data_ptr(), M, 3, 2, L, S, H, None, None, gridtype, align_corners, 0, 1, 0)
    return g_table, gw_s, gw_c


class _fused_field(Function):
    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, xyzs, dirs, embeddings, offsets, sigma_w, color_w, cfg):
        sigma, rgb, stash = field_forward(xyzs, dirs, embeddings, offsets, sigma_w, color_w, cfg)
        if stash is not None:
            ctx.save_for_backward(*stash["tensors"])
            ctx.cfg = stash["cfg"]
            ctx.sinks = stash["sinks"]
        return sigma, rgb

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, d_sigma, d_rgb):
        sink_t, sink_s, sink_c = ctx.sinks
        g_table, gw_s, gw_c = field_backward(ctx.saved_tensors, ctx.cfg, ctx.sinks, d_sigma, d_rgb)
        return (None, None, None if sink_t is not None else g_table, None, None if sink_s is not None else gw_s,
                None if sink_c is not None else gw_c, None)


def field_cfg(encoder, sigma_net, color_net, bound, training):
    if encoder.input_dim != 3 or encoder.level_dim != 2 or encoder.interp_id != 0:
        raise RuntimeError("fused_field: needs a 3-D, 2-feature, linearly interpolated GridEncoder")
    if sigma_net.hidden_dim != 64 or color_net.hidden_dim != 64 or color_net.input_dim != 32 or sigma_net.padded_output_dim != 16:
        raise RuntimeError("fused_field: needs the network_ff topology (64-wide FFMLPs, 16 SH + 15 geo + 1 pad)")
    if sigma_net.activation != 0 or color_net.activation != 0:
        raise RuntimeError("fused_field: ReLU networks only")
    return (float(bound), float(encoder.per_level_scale), int(encoder.base_resolution), encoder.gridtype_id,
            bool(encoder.align_corners), sigma_net.num_layers, color_net.num_layers, bool(training))


def fused_field(encoder, sigma_net, color_net, x, d, bound=1, training=None):
    """sigma [M] fp32, rgb [M,3] fp32 for sample positions x [M,3] (world) and directions d [M,3]."""
    if training is None:
        training = sigma_net.training and torch.is_grad_enabled()
    cfg = field_cfg(encoder, sigma_net, color_net, bound, training)
    return _fused_field.apply(x.view(-1, 3), d.view(-1, 3), encoder.embeddings, encoder.offsets, sigma_net.weights,
                              color_net.weights, cfg)
