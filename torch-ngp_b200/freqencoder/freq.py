"""freqencoder.freq — frequency (positional) encoding behind the reference's names (FreqEncoder, reference freqencoder/freq.py:56-81;
freq_encode :15-53).  The op is ngp_autograd.FreqEncodeFn (CUDA: csrc/freq.cu)."""
import torch.nn as nn

from ngp_autograd import freq_encode   # noqa: F401  (re-exported under the reference's name)


class FreqEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree = input_dim, degree
        self.output_dim = input_dim * (1 + 2 * degree)          # x itself, then a sin and a cos block per octave

    def __repr__(self):
        return f"FreqEncoder: input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def forward(self, inputs, **kwargs):
        """inputs [..., input_dim] -> [..., output_dim] (float32)"""
        flat = inputs.reshape(-1, self.input_dim)
        enc = freq_encode(flat, self.degree, self.output_dim)
        return enc.reshape(list(inputs.shape[:-1]) + [self.output_dim])
