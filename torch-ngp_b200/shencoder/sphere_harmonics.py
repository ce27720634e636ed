"""shencoder — drop-in for the reference's shencoder/sphere_harmonics.py (SHEncoder :61-87, _sh_encoder :14-54).
CUDA: csrc/sh.cu via the C-ABI (include/ngp_b200.h)."""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

import _ngp_b200 as _backend


class _sh_encoder(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)  # force float32 for better precision
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        # inputs: [B, 3] float in [-1, 1]; RETURN [B, degree^2] float
        _backend.require_cuda(inputs)
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        B, input_dim = inputs.shape
        output_dim = degree ** 2
        outputs = torch.empty(B, output_dim, dtype=inputs.dtype, device=inputs.device)
        dy_dx = torch.empty(B, input_dim * output_dim, dtype=inputs.dtype, device=inputs.device) if calc_grad_inputs else None
        _backend.call("ngp_sh_encode_forward", inputs.data_ptr(), outputs.data_ptr(), B, input_dim, degree,
                      _backend.ptr(dy_dx))
        ctx.save_for_backward(inputs, dy_dx)
        ctx.dims = [B, input_dim, degree]
        return outputs

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        inputs, dy_dx = ctx.saved_tensors
        if dy_dx is not None:
            grad = grad.contiguous().float()
            B, input_dim, degree = ctx.dims
            grad_inputs = torch.zeros_like(inputs)
            _backend.call("ngp_sh_encode_backward", grad.data_ptr(), inputs.data_ptr(), B, input_dim, degree,
                          dy_dx.data_ptr(), grad_inputs.data_ptr())
            return grad_inputs, None, None
        return None, None, None


sh_encode = _sh_encoder.apply


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim  # coord dims, must be 3
        self.degree = degree        # 1 ~ 8
        self.output_dim = degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert self.degree > 0 and self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        # inputs: [..., input_dim] in [-size, size]; return [..., degree^2]
        inputs = inputs / size
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = sh_encode(inputs, self.degree, inputs.requires_grad)
        return outputs.reshape(prefix_shape + [self.output_dim])
