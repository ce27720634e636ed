"""raymarching — drop-in for the reference's raymarching/raymarching.py.

Same public functions, positional signatures, dtypes and return shapes (reference lines in each
docstring); the CUDA comes from csrc/raymarch.cu through the C-ABI (include/ngp_b200.h).
"""
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

import _ngp_b200 as _backend

__all__ = ['near_far_from_aabb', 'sph_from_ray', 'morton3D', 'morton3D_invert', 'packbits',
           'march_rays_train', 'composite_rays_train', 'march_rays', 'composite_rays']


def _f32c(t):
    t = t.contiguous()
    return t if t.dtype == torch.float32 else t.float()


# ----------------------------------------
# utils
# ----------------------------------------

class _near_far_from_aabb(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        ''' near/far of each ray against an AABB (reference raymarching.py:19-49).
        rays_o/rays_d: float [N, 3]; aabb: float [6] (xmin, ymin, zmin, xmax, ymax, zmax).
        Returns nears, fars: float [N] (FLT_MAX for a miss). '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = _f32c(rays_o).view(-1, 3)
        rays_d = _f32c(rays_d).view(-1, 3)
        aabb = _f32c(aabb.to(rays_o.device))
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        _backend.call("ngp_near_far_from_aabb", rays_o.data_ptr(), rays_d.data_ptr(), aabb.data_ptr(), N,
                      float(min_near), nears.data_ptr(), fars.data_ptr())
        return nears, fars

near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, radius):
        ''' (theta, phi) in [-1, 1] where each ray leaves the background sphere (reference :52-80). '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = _f32c(rays_o).view(-1, 3)
        rays_d = _f32c(rays_d).view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        _backend.call("ngp_sph_from_ray", rays_o.data_ptr(), rays_d.data_ptr(), float(radius), N, coords.data_ptr())
        return coords

sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        ''' coords int32 [N, 3] in [0, 1024) -> Morton indices int32 [N] (reference :83-103). '''
        if not coords.is_cuda: coords = coords.cuda()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        coords = coords.int().contiguous()
        _backend.call("ngp_morton3D", coords.data_ptr(), N, indices.data_ptr())
        return indices

morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        ''' Morton indices int32 [N] -> coords int32 [N, 3] (reference :106-126). '''
        if not indices.is_cuda: indices = indices.cuda()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        indices = indices.int().contiguous()
        _backend.call("ngp_morton3D_invert", indices.data_ptr(), N, coords.data_ptr())
        return coords

morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, grid, thresh, bitfield=None):
        ''' density grid float [C, H^3] -> occupancy bitfield uint8 [C * H^3 / 8] (reference :129-153). '''
        if not grid.is_cuda: grid = grid.cuda()
        grid = _f32c(grid)
        C = grid.shape[0]
        H3 = grid.shape[1]
        N = C * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        _backend.call("ngp_packbits", grid.data_ptr(), N, float(thresh), bitfield.data_ptr())
        return bitfield

packbits = _packbits.apply

# ----------------------------------------
# train functions
# ----------------------------------------

class _march_rays_train(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        ''' march rays through the occupancy grid, emitting sample points (reference :161-235).
        Returns xyzs [M,3], dirs [M,3], deltas [M,2] and rays int32 [N,3] = (ray id, offset, count);
        M follows the reference's rules: N*max_steps on the slow path (then cut to the used, aligned
        count), mean_count rounded up to `align` otherwise. '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        if not density_bitfield.is_cuda: density_bitfield = density_bitfield.cuda()
        rays_o = _f32c(rays_o).view(-1, 3)
        rays_d = _f32c(rays_d).view(-1, 3)
        density_bitfield = density_bitfield.contiguous()
        nears = _f32c(nears)
        fars = _f32c(fars)

        N = rays_o.shape[0]
        M = N * max_steps
        fast = (not force_all_rays) and mean_count > 0
        if fast:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count

        dev = rays_o.device
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.rand(N, dtype=rays_o.dtype, device=dev) if perturb else torch.zeros(N, dtype=rays_o.dtype, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)

        if fast:
            xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
            dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
            deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
            _backend.call("ngp_march_rays_train", rays_o.data_ptr(), rays_d.data_ptr(), density_bitfield.data_ptr(),
                          float(bound), float(dt_gamma), int(max_steps), N, int(C), int(H), M, nears.data_ptr(),
                          fars.data_ptr(), xyzs.data_ptr(), dirs.data_ptr(), deltas.data_ptr(), rays.data_ptr(),
                          step_counter.data_ptr(), noises.data_ptr())
            return xyzs, dirs, deltas, rays

        # slow path (first epochs / force_all_rays).  The reference allocates and zero-fills
        # N*max_steps rows (21 GB at 640k rays) and trims after a D2H read of the counter.  Same
        # results with bounded memory: count first (M=0 writes no samples, only rays/counter), read
        # the total, then march again into exactly-sized buffers.  Offsets are reproducible because
        # slot reservation is deterministic per launch shape (warp-ordered) up to block order, so the
        # second launch re-reserves from a fresh counter and `rays` is taken from it.
        probe = torch.zeros(2, dtype=torch.int32, device=dev)
        _backend.call("ngp_march_rays_train", rays_o.data_ptr(), rays_d.data_ptr(), density_bitfield.data_ptr(),
                      float(bound), float(dt_gamma), int(max_steps), N, int(C), int(H), 0, nears.data_ptr(),
                      fars.data_ptr(), None, None, None, rays.data_ptr(), probe.data_ptr(), noises.data_ptr())
        m = int(probe[0].item())  # D2H sync, as in the reference (:224)
        if align > 0:
            m += align - m % align
        m = min(m, N * max_steps) if N * max_steps > 0 else m
        xyzs = torch.zeros(m, 3, dtype=rays_o.dtype, device=dev)
        dirs = torch.zeros(m, 3, dtype=rays_o.dtype, device=dev)
        deltas = torch.zeros(m, 2, dtype=rays_o.dtype, device=dev)
        _backend.call("ngp_march_rays_train", rays_o.data_ptr(), rays_d.data_ptr(), density_bitfield.data_ptr(),
                      float(bound), float(dt_gamma), int(max_steps), N, int(C), int(H), m, nears.data_ptr(),
                      fars.data_ptr(), xyzs.data_ptr(), dirs.data_ptr(), deltas.data_ptr(), rays.data_ptr(),
                      step_counter.data_ptr(), noises.data_ptr())
        return xyzs, dirs, deltas, rays

march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        ''' alpha-composite samples per ray (reference :238-291).
        sigmas [M], rgbs [M,3], deltas [M,2], rays int32 [N,3] -> weights_sum [N], depth [N], image [N,3]. '''
        sigmas = _f32c(sigmas)
        rgbs = _f32c(rgbs)
        deltas = _f32c(deltas)
        rays = rays.contiguous()
        M = sigmas.shape[0]
        N = rays.shape[0]
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        _backend.call("ngp_composite_rays_train_forward", sigmas.data_ptr(), rgbs.data_ptr(), deltas.data_ptr(),
                      rays.data_ptr(), M, N, float(T_thresh), weights_sum.data_ptr(), depth.data_ptr(), image.data_ptr())
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, N, T_thresh]
        return weights_sum, depth, image

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # NOTE: grad_depth is not propagated (reference :275)
        grad_weights_sum = _f32c(grad_weights_sum)
        grad_image = _f32c(grad_image)
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        _backend.call("ngp_composite_rays_train_backward", grad_weights_sum.data_ptr(), grad_image.data_ptr(),
                      sigmas.data_ptr(), rgbs.data_ptr(), deltas.data_ptr(), rays.data_ptr(), weights_sum.data_ptr(),
                      image.data_ptr(), M, N, float(T_thresh), grad_sigmas.data_ptr(), grad_rgbs.data_ptr())
        return grad_sigmas, grad_rgbs, None, None, None

composite_rays_train = _composite_rays_train.apply

# ----------------------------------------
# infer functions
# ----------------------------------------

class _march_rays(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
                align=-1, perturb=False, dt_gamma=0, max_steps=1024):
        ''' march each alive ray by up to n_step occupied samples (reference :297-348).
        Returns xyzs [M,3], dirs [M,3], deltas [M,2] with M = n_alive*n_step rounded up to `align`. '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = _f32c(rays_o).view(-1, 3)
        rays_d = _f32c(rays_d).view(-1, 3)
        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)
        dev = rays_o.device
        xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
        noises = torch.rand(n_alive, dtype=rays_o.dtype, device=dev) if perturb else torch.zeros(n_alive, dtype=rays_o.dtype, device=dev)
        _backend.call("ngp_march_rays", int(n_alive), int(n_step), rays_alive.data_ptr(), rays_t.data_ptr(),
                      rays_o.data_ptr(), rays_d.data_ptr(), float(bound), float(dt_gamma), int(max_steps), int(C),
                      int(H), density_bitfield.data_ptr(), near.data_ptr(), far.data_ptr(), xyzs.data_ptr(),
                      dirs.data_ptr(), deltas.data_ptr(), noises.data_ptr())
        return xyzs, dirs, deltas

march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)  # sigmas & rgbs arrive as half under autocast
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        ''' accumulate n_step samples into the running per-ray image / depth / weights_sum, in place;
        rays that terminate get rays_alive[n] = -1 (reference :351-373). '''
        sigmas = _f32c(sigmas)
        rgbs = _f32c(rgbs)
        _backend.call("ngp_composite_rays", int(n_alive), int(n_step), float(T_thresh), rays_alive.data_ptr(),
                      rays_t.data_ptr(), sigmas.data_ptr(), rgbs.data_ptr(), deltas.data_ptr(), weights_sum.data_ptr(),
                      depth.data_ptr(), image.data_ptr())
        return tuple()

composite_rays = _composite_rays.apply
