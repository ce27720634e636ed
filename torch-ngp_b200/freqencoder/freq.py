"""freqencoder — drop-in for the reference's freqencoder/freq.py (FreqEncoder :56-81, _freq_encoder :15-50).
CUDA: csrc/freq.cu via the C-ABI (include/ngp_b200.h).  No CPU path."""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

import _ngp_b200 as _backend


class _freq_encoder(Function):
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)  # force float32 for better precision
    def forward(ctx, inputs, degree, output_dim):
        # inputs: [B, input_dim] float; RETURN [B, input_dim + 2 * degree * input_dim] float
        _backend.require_cuda(inputs)
        inputs = inputs.contiguous().float()
        B, input_dim = inputs.shape
        outputs = torch.empty(B, output_dim, dtype=inputs.dtype, device=inputs.device)
        _backend.call("ngp_freq_encode_forward", inputs.data_ptr(), B, input_dim, degree, output_dim, outputs.data_ptr())
        ctx.save_for_backward(inputs, outputs)
        ctx.dims = [B, input_dim, degree, output_dim]
        return outputs

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        grad = grad.contiguous().float()
        inputs, outputs = ctx.saved_tensors
        B, input_dim, degree, output_dim = ctx.dims
        grad_inputs = torch.empty_like(inputs)
        _backend.call("ngp_freq_encode_backward", grad.data_ptr(), outputs.data_ptr(), B, input_dim, degree, output_dim,
                      grad_inputs.data_ptr())
        return grad_inputs, None, None


freq_encode = _freq_encoder.apply


class FreqEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = input_dim + input_dim * 2 * degree

    def __repr__(self):
        return f"FreqEncoder: input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def forward(self, inputs, **kwargs):
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = freq_encode(inputs, self.degree, self.output_dim)
        return outputs.reshape(prefix_shape + [self.output_dim])
