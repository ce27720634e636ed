"""gridencoder.grid — the multiresolution hash-grid encoder module behind the reference's name (GridEncoder, reference
gridencoder/grid.py:96-185; grid_encode :24-93): same constructor arguments, attributes and state-dict keys (`embeddings`
[entries, level_dim] fp32, `offsets` [levels+1] int32), so encoding.get_encoder('hashgrid' | 'tiledgrid') and trained
checkpoints work unchanged.  The op itself lives in ngp_autograd.GridEncodeFn (CUDA: csrc/grid.cu through the C ABI).

Invisible differences: features are produced directly as [B, levels*level_dim] (no [L,B,C] tensor + permute copy, grid.py:57,75)
and, when ngp_optim.FusedFieldOptimizer owns the parameters, the fp16 copy of the table used under autocast is the shadow that
optimizer's kernel keeps current instead of a per-forward cast (grid.py:43-44); without that owner it is re-cast every forward.
"""
import numpy as np
import torch
import torch.nn as nn

import _ngp_b200 as _backend
from ngp_autograd import grid_encode, _half_table, sync_half_table   # noqa: F401  (re-exported under the reference's names)

GRID_TYPES = {'hash': 0, 'tiled': 1}
INTERPOLATIONS = {'linear': 0, 'smoothstep': 1}


def level_table(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners):
    """First-entry offset of every level (+ the total): a level stores min(2^log2_hashmap_size, side^input_dim) entries, side =
    ceil(base * scale^l) (+1 unless align_corners), rounded up to a multiple of 8 — the arithmetic of reference grid.py:117-129."""
    cap = 2 ** log2_hashmap_size
    table = [0]
    for level in range(num_levels):
        # scalar arithmetic on purpose: ceil() sits on exact powers (16 * 2^(7/15)^15 = 2048), the rounding must match the reference's
        side = int(np.ceil(base_resolution * per_level_scale ** level)) + (0 if align_corners else 1)
        entries = min(cap, side ** input_dim)
        table.append(table[-1] + (entries + 7) // 8 * 8)
    return np.asarray(table, dtype=np.int32)


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype='hash', align_corners=False, interpolation='linear'):
        super().__init__()
        if desired_resolution is not None:
            # a target finest resolution overrides the growth factor
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution, self.log2_hashmap_size = per_level_scale, base_resolution, log2_hashmap_size
        self.gridtype, self.gridtype_id = gridtype, GRID_TYPES[gridtype]
        self.interpolation, self.interp_id = interpolation, INTERPOLATIONS[interpolation]
        self.align_corners = align_corners
        self.output_dim = num_levels * level_dim
        self.max_params = 2 ** log2_hashmap_size

        offsets = level_table(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners)
        self.register_buffer('offsets', torch.from_numpy(offsets))
        self.n_params = self.offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.reset_parameters()
        # load_state_dict copies into .data (no version bump): keep an owner-maintained fp16 shadow in step
        self.register_load_state_dict_post_hook(lambda module, incompatible: sync_half_table(module.embeddings))

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)
        sync_half_table(self.embeddings)     # a `.data` write: an owner-maintained fp16 shadow must follow it

    def __repr__(self):
        finest = int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {finest} per_level_scale={self.per_level_scale:.4f} "
                f"params={tuple(self.embeddings.shape)} gridtype={self.gridtype} align_corners={self.align_corners} "
                f"interpolation={self.interpolation}")

    def forward(self, inputs, bound=1):
        """inputs [..., input_dim] in [-bound, bound] -> [..., num_levels * level_dim].
        Where the fused encoder->MLP kernel reproduces this op exactly (ngp_lazy.defer_grid) the result is a deferred tensor: a
        drop-in FFMLP consumes it without the features ever reaching HBM, anything else materialises it through _forward_eager."""
        import ngp_lazy
        deferred = ngp_lazy.defer_grid(self, inputs, bound)
        return deferred if deferred is not None else self._forward_eager(inputs, bound)

    def _forward_eager(self, inputs, bound=1):
        unit = (inputs + bound) / (2 * bound)          # kept as a torch op so autograd w.r.t. the coordinates works
        lead = list(unit.shape[:-1])
        flat = unit.view(-1, self.input_dim)
        feats = grid_encode(flat, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution, flat.requires_grad,
                            self.gridtype_id, self.align_corners, self.interp_id)
        return feats.view(lead + [self.output_dim])

    @torch.amp.autocast('cuda', enabled=False)           # always in float precision
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        """adds weight * d(total variation)/d(table) into embeddings.grad at the cells of `inputs` (or B random points)"""
        if self.embeddings.grad is None:
            raise ValueError('grad is None, should be called after loss.backward() and before optimizer.step()!')
        if inputs is None:
            pts = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            pts = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
        pts = pts.contiguous().to(self.embeddings.dtype)
        _backend.call("ngp_grad_total_variation", pts.data_ptr(), self.embeddings.data_ptr(), self.embeddings.grad.data_ptr(),
                      self.offsets.data_ptr(), float(weight), pts.shape[0], self.input_dim, self.embeddings.shape[1],
                      self.offsets.shape[0] - 1, float(np.log2(self.per_level_scale)), self.base_resolution, self.gridtype_id,
                      int(self.align_corners), 0 if self.embeddings.dtype == torch.float32 else 1)
