"""raymarching.raymarching — occupancy-grid ray marching and volumetric compositing behind the reference's function names
(reference raymarching/raymarching.py: near_far_from_aabb :19-49, sph_from_ray :52-80, morton3D :83-103, morton3D_invert :106-126,
packbits :129-153, march_rays_train :161-235, composite_rays_train :238-291, march_rays :297-348, composite_rays :351-373).
nerf/renderer.py calls them positionally; parameter order, defaults and return shapes are the reference's.

Only composite_rays_train is differentiable, so it alone is an autograd Function; the other eight are plain functions that cast
their floating inputs to float32 (the reference's custom_fwd(cast_inputs=float32) contract) and hand raw pointers to the C ABI
(include/ngp_b200.h, CUDA in csrc/raymarch.cu) on the current stream.  Inputs on the CPU are moved to the GPU like the reference
does; there is no CPU implementation.
"""
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

import _ngp_b200 as _backend

__all__ = ['near_far_from_aabb', 'sph_from_ray', 'morton3D', 'morton3D_invert', 'packbits',
           'march_rays_train', 'composite_rays_train', 'march_rays', 'composite_rays']

_call = _backend.call


def _f32(t):
    """contiguous float32 view/copy of a tensor (no-op for what the renderer normally passes)"""
    t = t.contiguous()
    return t if t.dtype == torch.float32 else t.float()


def _gpu(t):
    return t if t.is_cuda else t.cuda()


def _ray_pair(rays_o, rays_d):
    return _f32(_gpu(rays_o)).view(-1, 3), _f32(_gpu(rays_d)).view(-1, 3)


def _sample_buffers(m, dev):
    """zero-initialised (xyzs [m,3], dirs [m,3], deltas [m,2]) — rows the marcher does not reach must read as zeros"""
    return (torch.zeros(m, 3, dtype=torch.float32, device=dev), torch.zeros(m, 3, dtype=torch.float32, device=dev),
            torch.zeros(m, 2, dtype=torch.float32, device=dev))


def _round_up(m, align):
    # the reference's rule, `m += align - m % align`: a multiple of `align` still grows by one full block
    return m + (align - m % align) if align > 0 else m


# ------------------------------------------------------------------------------------------------ utilities
@torch.no_grad()
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """entry / exit distance of each ray against the box aabb = (xmin, ymin, zmin, xmax, ymax, zmax); FLT_MAX pair for a miss.
    rays_o, rays_d [N,3] -> nears [N], fars [N]"""
    o, d = _ray_pair(rays_o, rays_d)
    box = _f32(aabb.to(o.device))
    n = o.shape[0]
    nears = torch.empty(n, dtype=torch.float32, device=o.device)
    fars = torch.empty_like(nears)
    _call("ngp_near_far_from_aabb", o.data_ptr(), d.data_ptr(), box.data_ptr(), n, float(min_near), nears.data_ptr(), fars.data_ptr())
    return nears, fars


@torch.no_grad()
def sph_from_ray(rays_o, rays_d, radius):
    """(theta, phi) in [-1, 1]^2 of the point where each ray leaves the background sphere -> [N,2]"""
    o, d = _ray_pair(rays_o, rays_d)
    n = o.shape[0]
    coords = torch.empty(n, 2, dtype=torch.float32, device=o.device)
    _call("ngp_sph_from_ray", o.data_ptr(), d.data_ptr(), float(radius), n, coords.data_ptr())
    return coords


@torch.no_grad()
def morton3D(coords):
    """cell coordinates int [N,3] in [0,1024) -> Morton (Z-order) index int32 [N]"""
    c = _gpu(coords).int().contiguous()
    idx = torch.empty(c.shape[0], dtype=torch.int32, device=c.device)
    _call("ngp_morton3D", c.data_ptr(), c.shape[0], idx.data_ptr())
    return idx


@torch.no_grad()
def morton3D_invert(indices):
    """Morton index int [N] -> cell coordinates int32 [N,3]"""
    i = _gpu(indices).int().contiguous()
    c = torch.empty(i.shape[0], 3, dtype=torch.int32, device=i.device)
    _call("ngp_morton3D_invert", i.data_ptr(), i.shape[0], c.data_ptr())
    return c


@torch.no_grad()
def packbits(grid, thresh, bitfield=None):
    """density grid float [C, H^3] -> occupancy bits uint8 [C*H^3/8] (bit k of byte n = cell 8n+k above thresh); written in place
    when `bitfield` is given (a slice view works: dnerf keeps one bitfield per time step)"""
    g = _f32(_gpu(grid))
    nbytes = g.shape[0] * g.shape[1] // 8
    if bitfield is None:
        bitfield = torch.empty(nbytes, dtype=torch.uint8, device=g.device)
    _call("ngp_packbits", g.data_ptr(), nbytes, float(thresh), bitfield.data_ptr())
    return bitfield


# ------------------------------------------------------------------------------------------------ training
@torch.no_grad()
def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False,
                     align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
    """March every ray through the occupancy grid and emit its sample points.
    Returns xyzs [M,3], dirs [M,3], deltas [M,2] (step size, distance to the previous sample) and rays int32 [N,3] =
    (ray id, first sample, sample count).  M follows the reference: with a sample budget (mean_count > 0 and not force_all_rays) M =
    mean_count rounded up to `align` and rays that do not fit are dropped; otherwise every sample is kept."""
    o, d = _ray_pair(rays_o, rays_d)
    bits = _gpu(density_bitfield).contiguous()
    near, far = _f32(nears), _f32(fars)
    n, dev = o.shape[0], o.device
    counter = step_counter if step_counter is not None else torch.zeros(2, dtype=torch.int32, device=dev)
    noise = torch.rand(n, dtype=torch.float32, device=dev) if perturb else torch.zeros(n, dtype=torch.float32, device=dev)
    rays = torch.empty(n, 3, dtype=torch.int32, device=dev)

    def launch(capacity, bufs, ctr):
        _call("ngp_march_rays_train", o.data_ptr(), d.data_ptr(), bits.data_ptr(), float(bound), float(dt_gamma), int(max_steps), n,
              int(C), int(H), capacity, near.data_ptr(), far.data_ptr(), *(_backend.ptr(b) for b in bufs), rays.data_ptr(),
              ctr.data_ptr(), noise.data_ptr())

    if not force_all_rays and mean_count > 0:
        bufs = _sample_buffers(_round_up(mean_count, align), dev)
        launch(bufs[0].shape[0], bufs, counter)
        return (*bufs, rays)

    # No budget yet (first epochs / force_all_rays).  The reference allocates and zero-fills N*max_steps rows (21 GB at 640 000 rays)
    # and trims after reading the counter back.  Same results with bounded memory: a counting pass (capacity 0 writes only `rays` and
    # the counter), one host read of the total — the reference synchronises here too (:224) — then the real pass into exact buffers.
    probe = torch.zeros(2, dtype=torch.int32, device=dev)
    launch(0, (None, None, None), probe)
    total = _round_up(int(probe[0].item()), align)
    if n * max_steps > 0:
        total = min(total, n * max_steps)
    bufs = _sample_buffers(total, dev)
    launch(total, bufs, counter)
    return (*bufs, rays)


class CompositeRaysTrainFn(Function):
    """per-ray alpha compositing of (sigma, rgb, delta) samples -> weights_sum [N], depth [N], image [N,3]; depth gets no gradient"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        s, c, dl, r = _f32(sigmas), _f32(rgbs), _f32(deltas), rays.contiguous()
        m, n = s.shape[0], r.shape[0]
        wsum = torch.empty(n, dtype=torch.float32, device=s.device)
        depth = torch.empty_like(wsum)
        image = torch.empty(n, 3, dtype=torch.float32, device=s.device)
        _call("ngp_composite_rays_train_forward", s.data_ptr(), c.data_ptr(), dl.data_ptr(), r.data_ptr(), m, n, float(T_thresh),
              wsum.data_ptr(), depth.data_ptr(), image.data_ptr())
        ctx.save_for_backward(s, c, dl, r, wsum, image)
        ctx.T_thresh = float(T_thresh)
        return wsum, depth, image

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        s, c, dl, r, wsum, image = ctx.saved_tensors
        g_w, g_i = _f32(grad_weights_sum), _f32(grad_image)
        d_sigma, d_rgb = torch.zeros_like(s), torch.zeros_like(c)
        _call("ngp_composite_rays_train_backward", g_w.data_ptr(), g_i.data_ptr(), s.data_ptr(), c.data_ptr(), dl.data_ptr(), r.data_ptr(),
              wsum.data_ptr(), image.data_ptr(), s.shape[0], r.shape[0], ctx.T_thresh, d_sigma.data_ptr(), d_rgb.data_ptr())
        return d_sigma, d_rgb, None, None, None


composite_rays_train = CompositeRaysTrainFn.apply


# ------------------------------------------------------------------------------------------------ inference
@torch.no_grad()
def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
               perturb=False, dt_gamma=0, max_steps=1024):
    """Advance each alive ray by up to n_step occupied samples starting at rays_t.
    Returns xyzs [M,3], dirs [M,3], deltas [M,2] with M = n_alive * n_step rounded up to `align` (unused rows stay zero)."""
    o, d = _ray_pair(rays_o, rays_d)
    dev = o.device
    bufs = _sample_buffers(_round_up(n_alive * n_step, align), dev)
    noise = (torch.rand(n_alive, dtype=torch.float32, device=dev) if perturb
             else torch.zeros(n_alive, dtype=torch.float32, device=dev))
    _call("ngp_march_rays", int(n_alive), int(n_step), rays_alive.data_ptr(), rays_t.data_ptr(), o.data_ptr(), d.data_ptr(),
          float(bound), float(dt_gamma), int(max_steps), int(C), int(H), density_bitfield.data_ptr(), near.data_ptr(), far.data_ptr(),
          bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), noise.data_ptr())
    return bufs


@torch.no_grad()
def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """Accumulate the n_step samples of each alive ray into weights_sum / depth / image IN PLACE, advance rays_t, and mark rays whose
    transmittance fell below T_thresh (or that ran out of samples) with rays_alive[n] = -1.  sigmas / rgbs may arrive as half."""
    s, c = _f32(sigmas), _f32(rgbs)
    _call("ngp_composite_rays", int(n_alive), int(n_step), float(T_thresh), rays_alive.data_ptr(), rays_t.data_ptr(), s.data_ptr(),
          c.data_ptr(), deltas.data_ptr(), weights_sum.data_ptr(), depth.data_ptr(), image.data_ptr())
    return tuple()
