"""gridencoder — drop-in for the reference's gridencoder/grid.py (same names, arguments, state-dict keys).

`GridEncoder` (reference grid.py:96-185) and `grid_encode` (grid.py:24-93) keep their signatures;
underneath, the CUDA comes from libngp_b200.so (csrc/grid.cu) through the C-ABI in include/ngp_b200.h.
Differences that are invisible to callers:
  * the kernel writes [B, L*C] directly (no [L,B,C] tensor + permute copy, grid.py:57,75),
  * the world->[0,1] affine map (grid.py:149) is kept as a torch op so autograd w.r.t. inputs works,
  * the fp16 shadow of the table used under autocast (grid.py:43-44) is cached per parameter version
    instead of being re-cast on every forward.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

import _ngp_b200 as _backend

_gridtype_to_id = {'hash': 0, 'tiled': 1}
_interp_to_id = {'linear': 0, 'smoothstep': 1}

def _half_table(embeddings):
    """fp16 shadow of the table, cached ON the parameter object and invalidated by its autograd version counter
    (optimizer steps bump it).  Keying by data_ptr would be wrong: freed tables get their address reused."""
    ver = embeddings._version
    hit = getattr(embeddings, "_ngp_half_shadow", None)
    if hit is not None and hit[0] == ver and hit[1].shape == embeddings.shape and hit[1].device == embeddings.device:
        return hit[1]
    half = embeddings.detach().to(torch.half)
    try:
        embeddings._ngp_half_shadow = (ver, half)
    except Exception:
        pass
    return half


class _grid_encode(Function):
    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0, align_corners=False, interpolation=0):
        # inputs: [B, D] float in [0, 1]; embeddings: [sO, C]; offsets: [L + 1] int32; RETURN [B, L*C]
        _backend.require_cuda(inputs, embeddings, offsets)
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            raise RuntimeError("grid_encode: inputs must be float32 (reference: inputs.data_ptr<float>())")
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)

        # autocast: half-precision table, float coordinates (reference grid.py:41-44)
        if torch.is_autocast_enabled('cuda') and C % 2 == 0:
            table = _half_table(embeddings)
        else:
            table = embeddings.detach().contiguous()
        if table.dtype not in (torch.float32, torch.float16):
            raise RuntimeError("grid_encode: embeddings must be float32 or float16")
        dtype = 1 if table.dtype == torch.float16 else 0

        outputs = torch.empty(B, L * C, device=inputs.device, dtype=table.dtype)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=table.dtype) if calc_grad_inputs else None

        _backend.call("ngp_grid_encode_forward", inputs.data_ptr(), table.data_ptr(), offsets.data_ptr(),
                      outputs.data_ptr(), B, D, C, L, S, H, _backend.ptr(dy_dx), gridtype, int(align_corners),
                      interpolation, dtype, 0)

        ctx.save_for_backward(inputs, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, H, gridtype, interpolation, dtype]
        ctx.align_corners = align_corners
        ctx.table_shape = table.shape
        ctx.table_dtype = table.dtype
        return outputs

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        inputs, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation, dtype = ctx.dims
        grad = grad.contiguous()
        if grad.dtype != ctx.table_dtype:
            grad = grad.to(ctx.table_dtype)

        grad_embeddings = torch.zeros(ctx.table_shape, device=grad.device, dtype=ctx.table_dtype)
        grad_inputs = torch.zeros_like(inputs, dtype=ctx.table_dtype) if dy_dx is not None else None

        _backend.call("ngp_grid_encode_backward", grad.data_ptr(), inputs.data_ptr(), None, offsets.data_ptr(),
                      grad_embeddings.data_ptr(), B, D, C, L, S, H, _backend.ptr(dy_dx), _backend.ptr(grad_inputs),
                      gridtype, int(ctx.align_corners), interpolation, dtype, 0)

        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False,
                 interpolation='linear'):
        super().__init__()

        # the finest resolution desired at the last level; if given it overrides per_level_scale
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))

        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners

        # level table (same arithmetic as reference grid.py:117-129: entries rounded up to a multiple of 8)
        offsets = []
        offset = 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            params_in_level = min(self.max_params, (resolution if align_corners else resolution + 1) ** input_dim)
            params_in_level = int(np.ceil(params_in_level / 8) * 8)
            offsets.append(offset)
            offset += params_in_level
        offsets.append(offset)
        offsets = torch.from_numpy(np.array(offsets, dtype=np.int32))
        self.register_buffer('offsets', offsets)

        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        std = 1e-4
        self.embeddings.data.uniform_(-std, std)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> "
                f"{int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} "
                f"gridtype={self.gridtype} align_corners={self.align_corners} interpolation={self.interpolation}")

    def forward(self, inputs, bound=1):
        # inputs: [..., input_dim] in [-bound, bound]; return [..., num_levels * level_dim]
        inputs = (inputs + bound) / (2 * bound)  # map to [0, 1]
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners, self.interp_id)
        return outputs.view(prefix_shape + [self.output_dim])

    # always run in float precision!
    @torch.amp.autocast('cuda', enabled=False)
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        D = self.input_dim
        C = self.embeddings.shape[1]
        L = self.offsets.shape[0] - 1
        S = float(np.log2(self.per_level_scale))
        H = self.base_resolution
        if inputs is None:
            inputs = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            inputs = (inputs + bound) / (2 * bound)
            inputs = inputs.view(-1, self.input_dim)
            B = inputs.shape[0]
        if self.embeddings.grad is None:
            raise ValueError('grad is None, should be called after loss.backward() and before optimizer.step()!')
        inputs = inputs.contiguous().to(self.embeddings.dtype)
        _backend.call("ngp_grad_total_variation", inputs.data_ptr(), self.embeddings.data_ptr(),
                      self.embeddings.grad.data_ptr(), self.offsets.data_ptr(), float(weight), B, D, C, L, S, H,
                      self.gridtype_id, int(self.align_corners), 0 if self.embeddings.dtype == torch.float32 else 1)
