"""ref_driver.py — drive the UNMODIFIED reference CUDA extensions (oracle/_ref/*.so, built by
oracle/build_ref.py from /root/reference/*/src) through their own pybind tables.

TEST INFRASTRUCTURE ONLY (GPU oracle "1" of SURVEY §8c + the "reference CUDA build" timing arm).
The reference's Python wrappers live in /root/reference and cannot travel to the GPU box, so each
function here issues the same native calls, with the same argument preparation, as the wrapper it
cites.  Nothing under torch-ngp_b200/ imports this module.
"""
import importlib.machinery
import importlib.util
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_mods = {}


def available(name=None):
    names = [name] if name else ["gridencoder", "shencoder", "raymarching", "ffmlp"]
    return all(os.path.exists(os.path.join(_HERE, "_ref", n, f"_ref_{n}.so")) for n in names)


def mod(name):
    if name not in _mods:
        path = os.path.join(_HERE, "_ref", name, f"_ref_{name}.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference exists")
        loader = importlib.machinery.ExtensionFileLoader(f"_ref_{name}", path)
        spec = importlib.util.spec_from_loader(f"_ref_{name}", loader)
        m = importlib.util.module_from_spec(spec)
        loader.exec_module(m)
        _mods[name] = m
    return _mods[name]


# ---- gridencoder/grid.py:24-90 ------------------------------------------------------------------
def grid_encode_forward(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                        gridtype=0, align_corners=False, interpolation=0):
    """Returns (outputs [B, L*C], dy_dx or None); `embeddings` is used in the dtype given (half = autocast path)."""
    inputs = inputs.contiguous()
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    S = np.log2(per_level_scale)
    H = base_resolution
    outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
    dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype) if calc_grad_inputs else None
    mod("gridencoder").grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype,
                                           align_corners, interpolation)
    return outputs.permute(1, 0, 2).reshape(B, L * C), dy_dx


def grid_encode_backward(grad, inputs, embeddings, offsets, per_level_scale, base_resolution, dy_dx=None, gridtype=0,
                         align_corners=False, interpolation=0):
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    S = np.log2(per_level_scale)
    H = base_resolution
    grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
    grad_embeddings = torch.zeros_like(embeddings)
    grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype) if dy_dx is not None else None
    mod("gridencoder").grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx,
                                            grad_inputs, gridtype, align_corners, interpolation)
    return grad_embeddings, grad_inputs


# ---- shencoder/sphere_harmonics.py:14-54 --------------------------------------------------------
def sh_encode_forward(inputs, degree, calc_grad_inputs=False):
    inputs = inputs.contiguous().float()
    B, D = inputs.shape
    outputs = torch.empty(B, degree ** 2, dtype=inputs.dtype, device=inputs.device)
    dy_dx = torch.empty(B, D * degree ** 2, dtype=inputs.dtype, device=inputs.device) if calc_grad_inputs else None
    mod("shencoder").sh_encode_forward(inputs, outputs, B, D, degree, dy_dx)
    return outputs, dy_dx


def sh_encode_backward(grad, inputs, degree, dy_dx):
    B, D = inputs.shape
    grad_inputs = torch.zeros_like(inputs)
    mod("shencoder").sh_encode_backward(grad.contiguous(), inputs, B, D, degree, dy_dx, grad_inputs)
    return grad_inputs


# ---- ffmlp/ffmlp.py:15-83 -----------------------------------------------------------------------
_splitk = {"n": 0}


def _ensure_splitk(n):
    if _splitk["n"] < n:
        mod("ffmlp").allocate_splitk(n)
        _splitk["n"] = n


def ffmlp_forward(inputs, weights, input_dim, output_dim, hidden_dim, num_layers, activation=0, output_activation=6,
                  inference=False):
    """inputs [B(%128==0), in] half, weights flat half.  Returns (outputs, forward_buffer or None)."""
    _ensure_splitk(num_layers + 1)
    B = inputs.shape[0]
    outputs = torch.empty(B, output_dim, device=inputs.device, dtype=inputs.dtype)
    if not inference:
        fb = torch.empty(num_layers, B, hidden_dim, device=inputs.device, dtype=inputs.dtype)
        mod("ffmlp").ffmlp_forward(inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation,
                                   output_activation, fb, outputs)
        return outputs, fb
    ib = torch.empty(B, hidden_dim, device=inputs.device, dtype=inputs.dtype)
    mod("ffmlp").ffmlp_inference(inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation,
                                 output_activation, ib, outputs)
    return outputs, None


def ffmlp_backward(grad, inputs, weights, forward_buffer, input_dim, output_dim, hidden_dim, num_layers, activation=0,
                   output_activation=6, calc_grad_inputs=True):
    _ensure_splitk(num_layers + 1)
    B = grad.shape[0]
    grad_inputs = torch.zeros_like(inputs) if calc_grad_inputs else torch.zeros(1, device=grad.device, dtype=grad.dtype)
    grad_weights = torch.zeros_like(weights)
    bb = torch.zeros(num_layers, B, hidden_dim, device=grad.device, dtype=grad.dtype)
    mod("ffmlp").ffmlp_backward(grad.contiguous(), inputs, weights, forward_buffer, B, input_dim, output_dim, hidden_dim,
                                num_layers, activation, output_activation, calc_grad_inputs, bb, grad_inputs, grad_weights)
    return (grad_inputs if calc_grad_inputs else None), grad_weights, bb


# ---- raymarching/raymarching.py -----------------------------------------------------------------
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    mod("raymarching").near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars)
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    N = rays_o.shape[0]
    coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
    mod("raymarching").sph_from_ray(rays_o, rays_d, radius, N, coords)
    return coords


def morton3D(coords):
    N = coords.shape[0]
    indices = torch.empty(N, dtype=torch.int32, device=coords.device)
    mod("raymarching").morton3D(coords.int(), N, indices)
    return indices


def morton3D_invert(indices):
    N = indices.shape[0]
    coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
    mod("raymarching").morton3D_invert(indices.int(), N, coords)
    return coords


def packbits(grid, thresh):
    grid = grid.contiguous()
    N = grid.shape[0] * grid.shape[1] // 8
    bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
    mod("raymarching").packbits(grid, N, thresh, bitfield)
    return bitfield


def march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, M, noises, dt_gamma=0.0, max_steps=1024,
                     counter=None):
    N = rays_o.shape[0]
    dev = rays_o.device
    xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev); deltas = torch.zeros(M, 2, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    if counter is None:
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
    mod("raymarching").march_rays_train(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, N, C, H, M, nears, fars,
                                        xyzs, dirs, deltas, rays, counter, noises)
    return xyzs, dirs, deltas, rays, counter


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    M, N = sigmas.shape[0], rays.shape[0]
    ws = torch.empty(N, device=sigmas.device); depth = torch.empty(N, device=sigmas.device)
    image = torch.empty(N, 3, device=sigmas.device)
    mod("raymarching").composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, ws, depth, image)
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, T_thresh=1e-4):
    M, N = sigmas.shape[0], rays.shape[0]
    gs = torch.zeros_like(sigmas); gc = torch.zeros_like(rgbs)
    mod("raymarching").composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, M, N,
                                                     T_thresh, gs, gc)
    return gs, gc


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises,
               align=-1, dt_gamma=0.0, max_steps=1024):
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    dev = rays_o.device
    xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev); deltas = torch.zeros(M, 2, device=dev)
    mod("raymarching").march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
                                  bitfield, nears, fars, xyzs, dirs, deltas, noises)
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
    mod("raymarching").composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum,
                                      depth, image)
