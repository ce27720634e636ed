"""oracle.py — numpy/ctypes front-end of the CPU oracle (oracle/ngp_oracle.c) + numpy MLP / SH oracles.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  Nothing under torch-ngp_b200/ imports this module.

Parity status: pinned by tests/golden/*.npz (outputs of the reference's own CUDA extensions, oracle/_ref,
recorded on a B200 by tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "ngp_oracle.c")
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

_c = ctypes
_u32, _i32, _f32, _vp = _c.c_uint32, _c.c_int, _c.c_float, _c.c_void_p


def build(force=False):
    """gcc -O2 -ffp-contract=off: no implicit FMA contraction; the C file calls fmaf() where nvcc fuses."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-fopenmp", "-o", _SO, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_grid_level_scale.restype = _f32
        _lib.oracle_grid_level_scale.argtypes = [_u32, _f32, _u32]
        _lib.oracle_morton3D.restype = _u32
        _lib.oracle_morton3D.argtypes = [_u32, _u32, _u32]
        _lib.oracle_morton3D_invert.restype = _u32
        _lib.oracle_morton3D_invert.argtypes = [_u32]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _c32(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


# ---------------------------------------------------------------- hash grid
def grid_offsets(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, align_corners=False):
    """Level offset table, restating gridencoder/grid.py:117-129.  Returns (offsets int32 [L+1], per_level_scale)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params = int(np.ceil(params / 8) * 8)
        offsets.append(offset)
        offset += params
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), per_level_scale


def grid_level_scales(L, S, H):
    return np.array([lib().oracle_grid_level_scale(l, np.float32(S), H) for l in range(L)], dtype=np.float32)


def grid_forward(inputs, table, offsets, S, H, gridtype=0, align_corners=False, interp=0, scales=None,
                 want_indices=False, want_dy_dx=False):
    """inputs [B,D] f32 in [0,1]; table [sO,C] f32|f16 -> out [B, L*C] (table dtype)."""
    inputs = _c32(inputs)
    B, D = inputs.shape
    assert table.dtype in (np.float32, np.float16)
    table = np.ascontiguousarray(table)
    C = table.shape[1]
    L = len(offsets) - 1
    dtype = 1 if table.dtype == np.float16 else 0
    out = np.zeros((B, L * C), dtype=table.dtype)
    idx = np.zeros((B, L, 1 << D), dtype=np.uint32) if want_indices else None
    dy = np.zeros((B, L * D * C), dtype=table.dtype) if want_dy_dx else None
    sc = None if scales is None else _c32(scales)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    lib().oracle_grid_forward(_p(inputs), _p(table), _p(offsets), _p(out), _u32(B), _u32(D), _u32(C), _u32(L),
                              _f32(S), _u32(H), _u32(gridtype), _i32(int(align_corners)), _u32(interp), _i32(dtype),
                              _p(sc), _p(idx), _p(dy))
    res = [out]
    if want_indices: res.append(idx)
    if want_dy_dx: res.append(dy)
    return res[0] if len(res) == 1 else tuple(res)


def grid_backward(grad, inputs, offsets, n_entries, C, S, H, gridtype=0, align_corners=False, interp=0, scales=None):
    """grad [B, L*C] (f32|f16) -> float64 grad table [n_entries, C] (order-independent scatter-add)."""
    inputs = _c32(inputs)
    B, D = inputs.shape
    grad = np.ascontiguousarray(grad)
    dtype = 1 if grad.dtype == np.float16 else 0
    L = len(offsets) - 1
    gt = np.zeros((n_entries, C), dtype=np.float64)
    sc = None if scales is None else _c32(scales)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    lib().oracle_grid_backward(_p(grad), _p(inputs), _p(offsets), _p(gt), _u32(B), _u32(D), _u32(C), _u32(L), _f32(S),
                               _u32(H), _u32(gridtype), _i32(int(align_corners)), _u32(interp), _i32(dtype), _p(sc))
    return gt


# ---------------------------------------------------------------- ray marching
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _c32(rays_o), _c32(rays_d), _c32(aabb)
    N = rays_o.shape[0]
    nears = np.zeros(N, np.float32); fars = np.zeros(N, np.float32)
    lib().oracle_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), _u32(N), _f32(min_near), _p(nears), _p(fars))
    return nears, fars


def morton3D(coords):
    coords = np.asarray(coords, dtype=np.uint32)
    return np.array([lib().oracle_morton3D(int(x), int(y), int(z)) for x, y, z in coords], dtype=np.int32)


def morton3D_invert(indices):
    out = np.zeros((len(indices), 3), np.int32)
    for i, v in enumerate(np.asarray(indices, dtype=np.uint32)):
        for k in range(3):
            out[i, k] = lib().oracle_morton3D_invert(int(v) >> k)
    return out


def packbits(grid, thresh):
    grid = _c32(grid).reshape(-1)
    N = grid.size // 8
    out = np.zeros(N, np.uint8)
    lib().oracle_packbits(_p(grid), _u32(N), _f32(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, C, H, M, nears, fars, noises):
    rays_o, rays_d, nears, fars, noises = map(_c32, (rays_o, rays_d, nears, fars, noises))
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    N = rays_o.shape[0]
    xyzs = np.zeros((M, 3), np.float32); dirs = np.zeros((M, 3), np.float32); deltas = np.zeros((M, 2), np.float32)
    rays = np.zeros((N, 3), np.int32); counter = np.zeros(2, np.int32)
    lib().oracle_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), _f32(bound), _f32(dt_gamma), _u32(max_steps),
                                  _u32(N), _u32(C), _u32(H), _u32(M), _p(nears), _p(fars), _p(xyzs), _p(dirs),
                                  _p(deltas), _p(rays), _p(counter), _p(noises))
    return xyzs, dirs, deltas, rays, counter


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, deltas = map(_c32, (sigmas, rgbs, deltas))
    rays = np.ascontiguousarray(rays, dtype=np.int32)
    M, N = sigmas.shape[0], rays.shape[0]
    ws = np.zeros(N, np.float32); depth = np.zeros(N, np.float32); image = np.zeros((N, 3), np.float32)
    lib().oracle_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), _u32(M), _u32(N),
                                              _f32(T_thresh), _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, T_thresh=1e-4):
    grad_ws, grad_image, sigmas, rgbs, deltas, ws, image = map(_c32, (grad_ws, grad_image, sigmas, rgbs, deltas, ws, image))
    rays = np.ascontiguousarray(rays, dtype=np.int32)
    M, N = sigmas.shape[0], rays.shape[0]
    gs = np.zeros(M, np.float32); gc = np.zeros((M, 3), np.float32)
    lib().oracle_composite_rays_train_backward(_p(grad_ws), _p(grad_image), _p(sigmas), _p(rgbs), _p(deltas), _p(rays),
                                               _p(ws), _p(image), _u32(M), _u32(N), _f32(T_thresh), _p(gs), _p(gc))
    return gs, gc


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, bitfield, nears,
               fars, noises, align=-1):
    rays_t, rays_o, rays_d, nears, fars, noises = map(_c32, (rays_t, rays_o, rays_d, nears, fars, noises))
    rays_alive = np.ascontiguousarray(rays_alive, dtype=np.int32)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs = np.zeros((M, 3), np.float32); dirs = np.zeros((M, 3), np.float32); deltas = np.zeros((M, 2), np.float32)
    lib().oracle_march_rays(_u32(n_alive), _u32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d), _f32(bound),
                            _f32(dt_gamma), _u32(max_steps), _u32(C), _u32(H), _p(bitfield), _p(nears), _p(fars),
                            _p(xyzs), _p(dirs), _p(deltas), _p(noises))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
    """In-place on copies; returns (rays_alive, rays_t, weights_sum, depth, image)."""
    rays_alive = np.array(rays_alive, dtype=np.int32)
    rays_t, weights_sum, depth, image = (np.array(a, dtype=np.float32) for a in (rays_t, weights_sum, depth, image))
    sigmas, rgbs, deltas = map(_c32, (sigmas, rgbs, deltas))
    lib().oracle_composite_rays(_u32(n_alive), _u32(n_step), _f32(T_thresh), _p(rays_alive), _p(rays_t), _p(sigmas),
                                _p(rgbs), _p(deltas), _p(weights_sum), _p(depth), _p(image))
    return rays_alive, rays_t, weights_sum, depth, image


# ---------------------------------------------------------------- MLP (numpy)
def _act(a, x):
    K = 10.0
    if a == 0: return np.maximum(x, 0)
    if a == 1: return np.exp(x)
    if a == 2: return np.sin(x)
    if a == 3: return 1 / (1 + np.exp(-x))
    if a == 4: t = x * K; return 0.5 * (t + np.sqrt(t * t + 4)) / K
    if a == 5: return np.log(np.exp(x * K) + 1) / K
    return x


def _act_bwd(a, g, f):
    K = 10.0
    if a == 0: return g * (f > 0)
    if a == 1: return g * f
    if a == 3: return g * f * (1 - f)
    if a == 4: y = f * K; return g * (y * y / (y * y + 1))
    if a == 5: return g * (1 - np.exp(-f * K))
    return g


def mlp_split_weights(weights, input_dim, hidden_dim, num_layers, output_dim=16):
    """Flat vector -> list of [out,in] matrices (layout of ffmlp/src/ffmlp.cu:631-634)."""
    mats, o = [], 0
    mats.append(weights[o:o + hidden_dim * input_dim].reshape(hidden_dim, input_dim)); o += hidden_dim * input_dim
    for _ in range(num_layers - 1):
        mats.append(weights[o:o + hidden_dim * hidden_dim].reshape(hidden_dim, hidden_dim)); o += hidden_dim * hidden_dim
    mats.append(weights[o:o + output_dim * hidden_dim].reshape(output_dim, hidden_dim))
    return mats


def mlp_forward(x, weights, input_dim, hidden_dim, num_layers, activation=0, output_dim=16):
    """Bias-free MLP as the reference defines it (testing/test_ffmlp.py:11-43 + ffmlp.cu:331-407):
    fp16 operands, products accumulated in fp32 here, every layer output rounded to fp16.
    Returns (y [B,16] f16, forward_buffer [num_layers,B,hidden] f16)."""
    mats = mlp_split_weights(np.asarray(weights, dtype=np.float16), input_dim, hidden_dim, num_layers, output_dim)
    h = np.asarray(x, dtype=np.float16)
    fwd = []
    for W in mats[:-1]:
        h = _act(activation, h.astype(np.float32) @ W.astype(np.float32).T).astype(np.float16)
        fwd.append(h)
    y = (h.astype(np.float32) @ mats[-1].astype(np.float32).T).astype(np.float16)
    return y, np.stack(fwd)


def mlp_backward(grad, x, weights, fwd, input_dim, hidden_dim, num_layers, activation=0, output_dim=16):
    """Returns (grad_inputs f16, grad_weights f32 flat, backward_buffer [num_layers,B,hidden] f16) following
    ffmlp.cu:749-895 (backward_buffer[j] = dL/dpre of hidden layer num_layers-1-j, each rounded to fp16)."""
    mats = mlp_split_weights(np.asarray(weights, dtype=np.float16), input_dim, hidden_dim, num_layers, output_dim)
    g = np.asarray(grad, dtype=np.float16).astype(np.float32)
    x = np.asarray(x, dtype=np.float16).astype(np.float32)
    gws = [None] * (num_layers + 1)
    bwd = []
    gws[num_layers] = g.T @ fwd[num_layers - 1].astype(np.float32)
    d = g
    for j in range(num_layers):
        layer = num_layers - 1 - j
        W = mats[layer + 1].astype(np.float32)
        d = _act_bwd(activation, d @ W, fwd[layer].astype(np.float32)).astype(np.float16)
        bwd.append(d)
        d = d.astype(np.float32)
        prev = fwd[layer - 1].astype(np.float32) if layer > 0 else x
        gws[layer] = d.T @ prev
    gi = (d @ mats[0].astype(np.float32)).astype(np.float16)
    return gi, np.concatenate([w.reshape(-1) for w in gws]).astype(np.float32), np.stack(bwd)


# ---------------------------------------------------------------- spherical harmonics (scipy, float64)
def sh_encode(dirs, degree):
    """Real SH basis with the reference's sign convention (shencoder.cu:50-121: Y_1^-1 = -c*y, Y_1^0 = c*z,
    Y_1^1 = -c*x ...), evaluated independently with scipy's complex harmonics.  NOTE the reference feeds the
    raw vector into polynomials that assume |d| = 1; this oracle is only valid for unit vectors."""
    from scipy.special import sph_harm_y
    d = np.asarray(dirs, dtype=np.float64)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    theta = np.arccos(np.clip(z, -1, 1))
    phi = np.arctan2(y, x)
    out = np.zeros((d.shape[0], degree * degree))
    for l in range(degree):
        for m in range(-l, l + 1):
            Y = sph_harm_y(l, abs(m), theta, phi)
            if m == 0:
                v = Y.real
            elif m > 0:
                v = np.sqrt(2) * Y.real
            else:
                v = np.sqrt(2) * Y.imag
            out[:, l * l + l + m] = v
    return out


# ---- frequency encoding (freqencoder/src/freqencoder.cu:28-104; same layout as the reference's pure-PyTorch FreqEncoder, encoding.py:5-42)
def freq_encode(x, degree):
    """[B,D] -> [B, D + 2*degree*D] = [x, sin(2^0 x), cos(2^0 x), ...] (float64 sin/cos rounded to fp32)."""
    x = _c32(x)
    out = [x]
    for f in range(degree):
        a = (x * np.float32(2.0 ** f)).astype(np.float64)
        out += [np.sin(a).astype(np.float32), np.cos(a).astype(np.float32)]
    return np.concatenate(out, axis=-1)


def freq_encode_backward(grad, outputs, D, degree):
    """:66-104 — grad_x = g_x + sum_f 2^f (g_sin * cos - g_cos * sin), sin/cos taken from the forward outputs."""
    g = np.asarray(grad, np.float64); y = np.asarray(outputs, np.float64)
    res = g[:, :D].copy()
    for f in range(degree):
        s0 = D + 2 * f * D
        res += (2.0 ** f) * (g[:, s0:s0 + D] * y[:, s0 + D:s0 + 2 * D] - g[:, s0 + D:s0 + 2 * D] * y[:, s0:s0 + D])
    return res.astype(np.float32)
