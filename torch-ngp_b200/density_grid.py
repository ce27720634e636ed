"""density_grid.py — occupancy-grid maintenance for the ray marcher.

Drop-in replacements for two methods of the reference's NeRFRenderer (nerf/renderer.py):

    mark_untrained_grid(self, poses, intrinsic, S=64)      renderer.py:380-442
    update_extra_state(self, decay=0.95, S=128)            renderer.py:445-538

Bind them onto the reference class (``NeRFRenderer.update_extra_state = density_grid.update_extra_state``) or call them
with any object that carries the reference's attributes (density_grid [C,H^3] f32 in Morton order, density_bitfield,
cascade, grid_size, bound, density_scale, density_thresh, iter_density, mean_density, step_counter, local_step, mean_count
and a density(x) method).  The work runs in the C-ABI kernels ngp_density_grid_* (include/ngp_b200.h): one kernel each for
frustum marking, ordered compaction of occupied cells, sample generation, and scatter / EMA-max / mean / threshold /
packbits with the threshold kept on the device.  Random numbers come from torch's generator, drawn in the reference's order
and shapes, so a full update consumes the RNG stream exactly as the reference does; the partial update replaces
torch.nonzero + randint(0, Nz) (a host sync per cascade) by a device-side pick unless exact_rng=True.

The density query between sampling and update is the fused encoder+MLP kernel (ngp_field_sigma_forward) when the model has
the network_ff topology, otherwise the model's own density().  There is no CPU path.
"""
import numpy as np
import torch

import _ngp_b200 as _backend


class _Workspace:
    """per-model device buffers (allocated once; the grid itself stays the model's registered buffer)"""

    def __init__(self, C, H, device):
        lib = _backend.load()
        H3 = H ** 3
        self.C, self.H = C, H
        self.tmp_grid = torch.empty(C, H3, dtype=torch.float32, device=device)
        self.occ_list = torch.empty(C, H3, dtype=torch.int32, device=device)
        self.occ_count = torch.zeros(C, dtype=torch.int32, device=device)
        self.occ_scratch = torch.empty(max(1, lib.ngp_density_grid_occupied_scratch_bytes(C, H) // 4), dtype=torch.int32, device=device)
        self.upd_scratch = torch.empty(max(1, lib.ngp_density_grid_update_scratch_bytes(C, H) // 8), dtype=torch.float64, device=device)
        self.state = torch.zeros(2, dtype=torch.float32, device=device)      # [mean_density, threshold used for the bitfield]


def _workspace(model):
    C, H, dev = int(model.cascade), int(model.grid_size), model.density_grid.device
    ws = getattr(model, "_ngp_dg_ws", None)
    if ws is None or ws.C != C or ws.H != H or ws.state.device != dev:
        ws = _Workspace(C, H, dev)
        object.__setattr__(model, "_ngp_dg_ws", ws)      # not a buffer / submodule: stays out of the state dict
    return ws


def _check(model):
    g = model.density_grid
    _backend.require_cuda(g, model.density_bitfield)
    if g.dtype != torch.float32 or not g.is_contiguous() or g.shape != (model.cascade, model.grid_size ** 3):
        raise RuntimeError("density_grid: expected a contiguous float32 [cascade, grid_size^3] buffer")
    return int(model.cascade), int(model.grid_size)


@torch.no_grad()
def mark_untrained_grid(self, poses, intrinsic, S=64, return_count=False):
    """renderer.py:380-442 in one kernel (S, the reference's chunk size, is accepted and ignored)."""
    if not getattr(self, "cuda_ray", True):
        return
    C, H = _check(self)
    if isinstance(poses, np.ndarray):
        poses = torch.from_numpy(poses)
    dev = self.density_grid.device
    poses = poses.to(device=dev, dtype=torch.float32).contiguous().view(-1, 4, 4)
    fx, fy, cx, cy = [float(v) for v in intrinsic]
    n_marked = torch.zeros(1, dtype=torch.int32, device=dev)
    count = torch.empty(C, H ** 3, dtype=torch.int32, device=dev) if return_count else None
    _backend.call("ngp_density_grid_mark_untrained", poses.data_ptr(), poses.shape[0], fx, fy, cx, cy, float(self.bound), C, H,
                  self.density_grid.data_ptr(), _backend.ptr(count), n_marked.data_ptr())
    print(f'[mark untrained grid] {int(n_marked.item())} from {H ** 3 * C}')
    return count


def fused_density_fn(model):
    """sigma(x) through ngp_field_sigma_forward (hash grid -> sigma MLP -> exp in one kernel) for the network_ff topology;
    None when the model is something else."""
    enc, net = getattr(model, "encoder", None), getattr(model, "sigma_net", None)
    try:
        from nerf_fused import field_cfg
        from gridencoder.grid import _half_table
        color = getattr(model, "color_net", None)
        cfg = field_cfg(enc, net, color, model.bound, False)
    except Exception:
        return None
    bound, pls, base, gridtype, align, nl_s = cfg[0], cfg[1], cfg[2], cfg[3], cfg[4], cfg[5]

    def fn(x):
        x01 = ((x.float() + bound) / (2 * bound)).contiguous()           # GridEncoder.forward's map (grid.py:149)
        M = x01.shape[0]
        sigma = torch.empty(M, dtype=torch.float32, device=x.device)
        table = _half_table(enc.embeddings)
        w = _half_table(net.weights).contiguous()      # owner-maintained fp16 copy when the fused optimizer holds one
        L = enc.offsets.shape[0] - 1
        _backend.call("ngp_field_sigma_forward", x01.data_ptr(), table.data_ptr(), enc.offsets.data_ptr(), L, float(np.log2(pls)),
                      int(base), gridtype, int(align), w.data_ptr(), nl_s, M, 0, None, None, None, sigma.data_ptr())
        return sigma
    return fn


def _density(model, density_fn, x):
    if density_fn is not None:
        return density_fn(x).reshape(-1).float().contiguous()
    return model.density(x)['sigma'].reshape(-1).detach().float().contiguous()


@torch.no_grad()
def update_extra_state(self, decay=0.95, S=128, density_fn="auto", exact_rng=False, sync=True):
    """renderer.py:445-538.  density_fn: "auto" (fused kernel if the topology allows, else self.density), None (always
    self.density) or a callable x[M,3] -> sigma[M].  exact_rng=True draws the occupied-cell picks with
    torch.randint(0, Nz) like the reference (one host sync per cascade).  sync=False leaves mean_density / mean_count
    untouched on the host (the threshold never leaves the device); the caller reads ws.state when it wants them."""
    if not getattr(self, "cuda_ray", True):
        return
    C, H = _check(self)
    ws = _workspace(self)
    dev = self.density_grid.device
    H3 = H ** 3
    if density_fn == "auto":
        density_fn = fused_density_fn(self)
    bound = float(self.bound)
    if self.iter_density < 16:
        # full update: every cell, jittered (renderer.py:456-483); one torch.rand per cascade, as the reference draws them
        noise = torch.stack([torch.rand(H3, 3, device=dev) for _ in range(C)])
        xyzs = torch.empty(C, H3, 3, dtype=torch.float32, device=dev)
        _backend.call("ngp_density_grid_sample_full", C, H, bound, noise.data_ptr(), xyzs.data_ptr())
        indices, N = None, H3
    else:
        # partial update: H^3/4 uniform cells + H^3/4 occupied cells per cascade (renderer.py:487-509)
        N = H3 // 4
        _backend.call("ngp_density_grid_occupied", self.density_grid.data_ptr(), C, H, ws.occ_list.data_ptr(),
                      ws.occ_count.data_ptr(), ws.occ_scratch.data_ptr())
        coords, picks, noise = [], [], []
        for cas in range(C):
            coords.append(torch.randint(0, H, (N, 3), device=dev))
            if exact_rng:
                nz = int(ws.occ_count[cas].item())
                if nz == 0:
                    raise RuntimeError("update_extra_state: no occupied cell to sample from (the reference fails here too)")
                picks.append(torch.randint(0, nz, [N], dtype=torch.long, device=dev))
            else:
                picks.append(torch.rand(N, device=dev))
            noise.append(torch.rand(2 * N, 3, device=dev))
        coords = torch.stack(coords).int().contiguous()
        picks, noise = torch.stack(picks).contiguous(), torch.stack(noise)
        xyzs = torch.empty(C, 2 * N, 3, dtype=torch.float32, device=dev)
        indices = torch.empty(C, 2 * N, dtype=torch.int32, device=dev)
        _backend.call("ngp_density_grid_sample_partial", C, H, bound, N, coords.data_ptr(),
                      picks.data_ptr() if exact_rng else None, None if exact_rng else picks.data_ptr(), ws.occ_list.data_ptr(),
                      ws.occ_count.data_ptr(), noise.data_ptr(), xyzs.data_ptr(), indices.data_ptr())
        N = 2 * N
    sigmas = _density(self, density_fn, xyzs.view(-1, 3))
    _backend.call("ngp_density_grid_update", self.density_grid.data_ptr(), ws.tmp_grid.data_ptr(), _backend.ptr(indices),
                  sigmas.data_ptr(), N, float(self.density_scale), float(decay), float(self.density_thresh), C, H,
                  self.density_bitfield.data_ptr(), ws.state.data_ptr(), ws.upd_scratch.data_ptr())
    self.iter_density += 1
    # step counter (renderer.py:532-536) — folded into the one host read of this call
    total_step = min(16, self.local_step)
    if sync:
        vals = torch.cat([ws.state[:1].double(), self.step_counter[:max(total_step, 1), 0].sum().double().view(1)]).tolist()
        self.mean_density = vals[0]
        if total_step > 0:
            self.mean_count = int(vals[1] / total_step)
    self.local_step = 0


def install(renderer_cls):
    """Rebind the two maintenance methods of a reference-shaped renderer class."""
    renderer_cls.mark_untrained_grid = mark_untrained_grid
    renderer_cls.update_extra_state = update_extra_state
    return renderer_cls
