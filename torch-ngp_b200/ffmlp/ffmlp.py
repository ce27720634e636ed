"""ffmlp.ffmlp — the fully fused MLP module behind the reference's name (FFMLP, reference ffmlp/ffmlp.py:99-168;
ffmlp_forward :15-86).  Same constructor, same parameter (`weights`: one flat fp32 vector holding the [out,in] row-major matrices
back to back, ffmlp.cu:631-634), same initialisation (global seed 42, U(+-sqrt(3/hidden)), ffmlp.py:141-144), output padded to 16
columns (ffmlp.py:118).  The op is ngp_autograd.FFMLPFn over the tcgen05/TMEM kernels of csrc/ffmlp.cu.

The reference pads the batch to a multiple of 128 with a cat-copy of the input (ffmlp.py:157-159); the kernels here mask the ragged
last tile instead, so no copy is made and results are identical.  (The stray `from turtle import ...` of ffmlp.py:2 is not reproduced.)
"""
import math

import torch
import torch.nn as nn

import _ngp_b200 as _backend
from ngp_autograd import ffmlp_forward   # noqa: F401  (re-exported under the reference's name)

ACTIVATION_IDS = {'relu': 0, 'exponential': 1, 'sine': 2, 'sigmoid': 3, 'squareplus': 4, 'softplus': 5}
ACT_NONE = 6


def convert_activation(act):
    """name -> the integer code of ffmlp.h's Activation enum (anything unknown means 'none')."""
    return ACTIVATION_IDS.get(act, ACT_NONE)


class FFMLP(nn.Module):
    tensorcore_width = 16

    def __init__(self, input_dim, output_dim, hidden_dim, num_layers, activation='relu'):
        super().__init__()
        # the reference's constraints (ffmlp.py:112-115); AssertionError like there
        assert hidden_dim in (16, 32, 64, 128, 256), f"FFMLP: hidden_dim {hidden_dim} not in (16, 32, 64, 128, 256)"
        assert input_dim > 0 and input_dim % 16 == 0, f"FFMLP: input_dim {input_dim} must be a positive multiple of 16"
        assert output_dim <= 16, f"FFMLP: output_dim {output_dim} > 16 is not supported"
        assert num_layers >= 2, f"FFMLP: num_layers {num_layers} < 2 (at least 3 matmuls)"
        # capability of THIS build, stated at construction instead of at the first forward (csrc/ffmlp.cu check_cfg): the tcgen05 kernels
        # are instantiated for 64 hidden units, at most 64 inputs and at most 8 hidden layers — what torch-ngp's NeRF / SDF networks use
        if hidden_dim != 64 or input_dim > 64 or num_layers > 8:
            raise NotImplementedError(f"FFMLP (B200 build): hidden_dim must be 64, input_dim <= 64, num_layers <= 8 "
                                      f"(got hidden_dim={hidden_dim}, input_dim={input_dim}, num_layers={num_layers}); the reference "
                                      f"additionally offers hidden widths 16/32/128/256 (ffmlp/src/ffmlp.cu:653-657)")
        self.input_dim, self.output_dim, self.hidden_dim, self.num_layers = input_dim, output_dim, hidden_dim, num_layers
        self.activation = convert_activation(activation)
        self.output_activation = ACT_NONE                          # the reference supports none either (ffmlp.py:108)
        self.padded_output_dim = -(-output_dim // 16) * 16
        self.num_parameters = hidden_dim * (input_dim + hidden_dim * (num_layers - 1) + self.padded_output_dim)
        self.weights = nn.Parameter(torch.zeros(self.num_parameters))
        self.reset_parameters()
        if torch.cuda.is_available():
            # the reference creates process-global split-K streams here (ffmlp.py:126); an ABI no-op in this library
            _backend.load().ngp_ffmlp_allocate_splitk(num_layers + 1)

    def cleanup(self):
        _backend.load().ngp_ffmlp_free_splitk()

    def reset_parameters(self):
        torch.manual_seed(42)      # reference behaviour: every FFMLP construction reseeds the GLOBAL generator
        bound = math.sqrt(3 / self.hidden_dim)
        self.weights.data.uniform_(-bound, bound)

    def __repr__(self):
        return (f"FFMLP: input_dim={self.input_dim} output_dim={self.output_dim} hidden_dim={self.hidden_dim} "
                f"num_layers={self.num_layers} activation={self.activation}")

    def forward(self, inputs):
        """inputs [B, input_dim] -> [B, output_dim] (half under autocast).  A deferred GridEncoder / SH-concat input (ngp_lazy) is
        consumed by the fused encoder->MLP kernel; the result is the same tensor the two separate ops would give."""
        import ngp_lazy
        if isinstance(inputs, ngp_lazy.Deferred):
            fused = None
            if isinstance(inputs, ngp_lazy.DeferredGridFeatures):
                fused = ngp_lazy.grid_mlp(inputs, self)
            elif isinstance(inputs, ngp_lazy.DeferredColorInput):
                fused = ngp_lazy.color_mlp(inputs, self)
            if fused is not None:
                return fused
            inputs = inputs.materialize()
        y = ffmlp_forward(inputs, self.weights, self.input_dim, self.padded_output_dim, self.hidden_dim, self.num_layers,
                          self.activation, self.output_activation, not self.training, inputs.requires_grad)
        return y if self.padded_output_dim == self.output_dim else y[:, :self.output_dim]
