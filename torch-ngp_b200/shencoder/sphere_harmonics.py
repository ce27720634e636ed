"""shencoder.sphere_harmonics — real spherical-harmonics direction encoding behind the reference's names (SHEncoder, reference
shencoder/sphere_harmonics.py:61-87; sh_encode :14-57).  The op is ngp_autograd.SHEncodeFn (CUDA: csrc/sh.cu)."""
import torch.nn as nn

from ngp_autograd import sh_encode   # noqa: F401  (re-exported under the reference's name)


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        assert input_dim == 3, "SHEncoder: directions are 3-D"
        assert 1 <= degree <= 8, "SHEncoder: degree must be in [1, 8]"
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree ** 2

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        """inputs [..., 3] in [-size, size] -> [..., degree^2] (float32).  Degree-4 encodings of [M,3] directions are returned as a
        deferred tensor (ngp_lazy): `torch.cat([sh, geo_feat, pad])` followed by a drop-in FFMLP then runs as one fused kernel; any
        other use materialises it through _forward_eager."""
        import ngp_lazy
        deferred = ngp_lazy.defer_sh(self, inputs, size)
        return deferred if deferred is not None else self._forward_eager(inputs, size)

    def _forward_eager(self, inputs, size=1):
        scaled = inputs / size
        flat = scaled.reshape(-1, self.input_dim)
        basis = sh_encode(flat, self.degree, flat.requires_grad)
        return basis.reshape(list(scaled.shape[:-1]) + [self.output_dim])
