"""GPU parity: spherical-harmonics encoder (csrc/sh.cu)."""
import numpy as np
import pytest
import torch

from util import gen, rel_err

pytestmark = pytest.mark.gpu


def _dirs(n, seed=7):
    d = torch.randn(n, 3, generator=gen(seed))
    return d / d.norm(dim=-1, keepdim=True)


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7, 8])
def test_forward_vs_scipy(degree):
    """fp32 recurrence vs float64 scipy harmonics: 1e-5 absolute (values are O(1))."""
    from oracle import oracle as O
    from shencoder import SHEncoder
    d = _dirs(3000)
    out = SHEncoder(degree=degree).cuda()(d.cuda())
    assert out.shape == (3000, degree * degree) and out.dtype == torch.float32
    ref = O.sh_encode(d.numpy(), degree)
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-5


def test_input_gradient():
    """Analytic Jacobian (dy_dx) vs central finite differences of the polynomial extension."""
    from shencoder import SHEncoder
    enc = SHEncoder(degree=4).cuda()
    d = _dirs(64).double()
    x = d.float().cuda().requires_grad_(True)
    g = torch.randn(64, 16, generator=gen(8)).cuda()
    enc(x).backward(g)
    eps = 1e-3
    num = torch.zeros(64, 3)
    for k in range(3):
        dp = d.clone(); dp[:, k] += eps
        dm = d.clone(); dm[:, k] -= eps
        fp = enc(dp.float().cuda()).double().cpu(); fm = enc(dm.float().cuda()).double().cpu()
        num[:, k] = (((fp - fm) / (2 * eps)) * g.double().cpu()).sum(-1)
    assert rel_err(x.grad.cpu().numpy(), num.numpy()) < 2e-3


def test_vs_reference_extension():
    from oracle import ref_driver as R
    if not R.available("shencoder"):
        pytest.skip("oracle/_ref/shencoder not built")
    import _ngp_b200 as nb
    d = (_dirs(100000) * (0.5 + torch.rand(100000, 1, generator=gen(9)))).cuda()   # also non-unit vectors
    for degree in (1, 4, 8):
        ref, ref_j = R.sh_encode_forward(d, degree, True)
        out = torch.empty_like(ref); jac = torch.empty_like(ref_j)
        nb.call("ngp_sh_encode_forward", d.data_ptr(), out.data_ptr(), d.shape[0], 3, degree, jac.data_ptr())
        assert (out - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
        assert (jac - ref_j).abs().max().item() < 2e-5 * max(1.0, ref_j.abs().max().item())
