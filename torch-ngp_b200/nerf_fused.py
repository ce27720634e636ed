"""nerf_fused.py — fused evaluation of the NeRF field (hash grid -> sigma MLP -> trunc_exp ; SH (+) geo -> color MLP -> sigmoid) on
top of the fused C-ABI extensions (include/ngp_b200.h, ngp_field_*).

Same mathematics, parameters and rounding points as the module-by-module path of nerf/network_ff.py:51-74 (reference) — encoder
features bit-identical, MLP outputs identical (same kernels), sigma/rgb within 1 ulp of torch.exp / torch.sigmoid — but the encoder
output feeds the tensor-core MLP from shared memory, SH + concat + casts happen in the color kernel's input staging, and the backward
runs sigmoid/cat/trunc_exp gradients inside the MLP backward.

Pipelining: the batch can be split into row chunks.  The forward then runs the color net of chunk k (HBM-write bound) on a side stream
while the gather-bound encoder+sigma kernel works on chunk k+1; the backward runs the hash-table scatter of chunk k (bound by the
SM's reduction-issue rate, no shared memory) on the side stream underneath the latency-bound tensor-core MLP backward kernels of
chunk k+1.  Chunks accumulate their weight gradients into one fp32 workspace (NGP_WGRAD_* flags); the table gradient is a
scatter-add, so chunk order does not matter.
"""
import numpy as np
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

import _ngp_b200 as _backend
from ngp_autograd import _half_table, _half_param

WGRAD_ACCUMULATE, WGRAD_NO_FINALIZE = 1, 2
DEFAULT_CHUNKS = 1          # autograd path (fused_field); the step driver passes its own chunk count


def chunk_ranges(M, chunks):
    """Row ranges [(r0, rows)] of `chunks` near-equal pieces whose starts are multiples of the 128-row MLP tile."""
    chunks = max(1, min(int(chunks), (M + 127) // 128))
    per = ((M + chunks - 1) // chunks + 127) // 128 * 128
    out = []
    r0 = 0
    while r0 < M:
        out.append((r0, min(per, M - r0)))
        r0 += per
    return out


def field_forward(xyzs, dirs, embeddings, offsets, sigma_w, color_w, cfg, chunks=1, side=None):
    """Raw (autograd-free) fused field evaluation.  Returns sigma [M] f32, rgb [M,3] f32 and the backward stash (or None).
    chunks > 1 with a side stream pipelines color(k) under sigma(k+1); the caller's stream has joined the side stream on return."""
    bound, pls, H, gridtype, align_corners, nl_s, nl_c, training = cfg
    _backend.require_cuda(xyzs, dirs, embeddings, sigma_w, color_w)
    x01 = ((xyzs.float() + bound) / (2 * bound)).contiguous()       # GridEncoder.forward's affine map (grid.py:149)
    dirs = dirs.float().contiguous()
    table = _half_table(embeddings)
    ws = _half_param(sigma_w)          # the optimizer's fp16 operand copy when it owns one, else a cast (ffmlp.py:18)
    wc = _half_param(color_w)
    M = x01.shape[0]
    L = offsets.shape[0] - 1
    S = float(np.log2(pls))
    dev = x01.device
    h = torch.empty(M, 16, dtype=torch.half, device=dev)
    sigma = torch.empty(M, dtype=torch.float32, device=dev)
    rgb = torch.empty(M, 3, dtype=torch.float32, device=dev)
    ranges = chunk_ranges(M, chunks) if M > 0 else []
    feat = torch.empty(M, 2 * L, dtype=torch.half, device=dev) if training else None
    fb_s = [torch.empty(nl_s, rows, 64, dtype=torch.half, device=dev) for _, rows in ranges] if training else []
    fb_c = [torch.empty(nl_c, rows, 64, dtype=torch.half, device=dev) for _, rows in ranges] if training else []
    main = torch.cuda.current_stream()
    piped = side is not None and len(ranges) > 1
    if piped:
        side.wait_stream(main)
    for k, (r0, rows) in enumerate(ranges):
        _backend.call("ngp_field_sigma_forward", x01.data_ptr() + 12 * r0, table.data_ptr(), offsets.data_ptr(), L, S, int(H),
                      gridtype, int(align_corners), ws.data_ptr(), nl_s, rows, int(training),
                      feat.data_ptr() + 4 * L * r0 if training else None, fb_s[k].data_ptr() if training else None,
                      h.data_ptr() + 32 * r0, sigma.data_ptr() + 4 * r0)
        if piped:
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
        with torch.cuda.stream(side if piped else main):
            _backend.call("ngp_field_color_forward", dirs.data_ptr() + 12 * r0, h.data_ptr() + 32 * r0, wc.data_ptr(), nl_c, rows,
                          int(training), fb_c[k].data_ptr() if training else None, rgb.data_ptr() + 12 * r0)
    if piped:
        main.wait_stream(side)
    stash = None
    if training:
        stash = dict(tensors=(x01, dirs, offsets, ws, wc, feat, h, rgb, *fb_s, *fb_c),
                     cfg=(L, S, int(H), gridtype, int(align_corners), nl_s, nl_c, M, tuple(table.shape), tuple(ranges)),
                     # optional fp16 gradient sinks installed by ngp_optim.FusedFieldOptimizer (bypass fp32 .grad accumulation)
                     sinks=tuple(getattr(p, "_ngp_grad_sink", None) for p in (embeddings, sigma_w, color_w)))
    return sigma, rgb, stash


_WGRAD_WS = {}


def _wgrad_workspace(dev, n):
    key = (str(dev), int(n))
    w = _WGRAD_WS.get(key)
    if w is None:
        w = _WGRAD_WS[key] = torch.zeros(n, dtype=torch.float32, device=dev)
    return w


def field_backward(tensors, cfg, sinks, d_sigma, d_rgb, side=None):
    """Raw backward of field_forward: returns (g_table, gw_sigma, gw_color) — the sinks themselves when installed.
    With a side stream and more than one chunk the table scatter of chunk k runs under the MLP backward kernels of chunk k+1."""
    L, S, H, gridtype, align_corners, nl_s, nl_c, M, table_shape, ranges = cfg
    x01, dirs, offsets, ws, wc, feat, h, rgb = tensors[:8]
    nck = len(ranges)
    fb_s, fb_c = tensors[8:8 + nck], tensors[8 + nck:8 + 2 * nck]
    dev = x01.device
    d_sigma = d_sigma.float().contiguous()
    d_rgb = d_rgb.float().contiguous()
    lib = _backend.load()
    sink_t, sink_s, sink_c = sinks
    gw_c = sink_c if sink_c is not None else torch.empty_like(wc)
    gw_s = sink_s if sink_s is not None else torch.empty_like(ws)
    # a sink is kept zeroed by the optimizer kernel; otherwise a fresh zero table as in grid.py:77
    g_table = sink_t if sink_t is not None else torch.zeros(table_shape, dtype=torch.half, device=dev)
    nb_c = lib.ngp_ffmlp_backward_workspace_bytes(M, 32, 16, 64, nl_c)
    nb_s = lib.ngp_ffmlp_backward_workspace_bytes(M, 2 * L, 16, 64, nl_s)
    # one persistent fp32 workspace [sigma | color] (the order of the weights in the optimizer's flat bucket), zero on creation and
    # cleared again by every finalize: no memset per step
    wk = _wgrad_workspace(dev, (nb_s + nb_c) // 4)
    wk_s, wk_c = wk[:nb_s // 4], wk[nb_s // 4:]
    dys = torch.empty(M, 16, dtype=torch.half, device=dev)          # dL/d(sigma-net output)
    d_feat = torch.empty(M, 2 * L, dtype=torch.half, device=dev)    # dL/d(encoder features)
    main = torch.cuda.current_stream()
    piped = side is not None and nck > 1
    if piped:
        side.wait_stream(main)
    flags = WGRAD_ACCUMULATE | WGRAD_NO_FINALIZE
    for k, (r0, rows) in enumerate(ranges):
        # color net (+ sigmoid, cat, trunc_exp gradients) -> dL/d(sigma-net output)
        _backend.call("ngp_field_color_backward_ex", d_rgb.data_ptr() + 12 * r0, rgb.data_ptr() + 12 * r0, None,
                      d_sigma.data_ptr() + 4 * r0, h.data_ptr() + 32 * r0, dirs.data_ptr() + 12 * r0, None, wc.data_ptr(),
                      fb_c[k].data_ptr(), nl_c, rows, dys.data_ptr() + 32 * r0, None, wk_c.data_ptr(), nb_c, flags)
        # sigma net -> dL/d(features)
        _backend.call("ngp_ffmlp_backward_ex", dys.data_ptr() + 32 * r0, feat.data_ptr() + 4 * L * r0, ws.data_ptr(), fb_s[k].data_ptr(),
                      rows, 2 * L, 16, 64, nl_s, 0, 6, 1, None, d_feat.data_ptr() + 4 * L * r0, None, wk_s.data_ptr(), nb_s, flags)
        if piped:
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
        # hash-table scatter-add
        with torch.cuda.stream(side if piped else main):
            _backend.call("ngp_grid_encode_backward", d_feat.data_ptr() + 4 * L * r0, x01.data_ptr() + 12 * r0, None, offsets.data_ptr(),
                          g_table.data_ptr(), rows, 3, 2, L, S, H, None, None, gridtype, align_corners, 0, 1, 0)
    if gw_c.data_ptr() == gw_s.data_ptr() + nb_s // 2:       # adjacent in the flat bucket: one launch converts both
        _backend.call("ngp_ffmlp_wgrad_finalize", wk.data_ptr(), gw_s.data_ptr(), (nb_s + nb_c) // 4, -1)
    else:
        _backend.call("ngp_ffmlp_wgrad_finalize", wk_s.data_ptr(), gw_s.data_ptr(), nb_s // 4, -1)
        _backend.call("ngp_ffmlp_wgrad_finalize", wk_c.data_ptr(), gw_c.data_ptr(), nb_c // 4, -1)
    if piped:
        main.wait_stream(side)
    return g_table, gw_s, gw_c


class _fused_field(Function):
    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, xyzs, dirs, embeddings, offsets, sigma_w, color_w, cfg):
        sigma, rgb, stash = field_forward(xyzs, dirs, embeddings, offsets, sigma_w, color_w, cfg, chunks=DEFAULT_CHUNKS)
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
