"""Frequency encoding: oracle vs the reference's pure-PyTorch FreqEncoder (tests/golden/freq.npz), module contract on CPU, and
(gpu) ngp_freq_encode_* vs both.  The CUDA kernel uses __sinf like the reference's extension: absolute error grows with the
argument (range reduction in fp32), hence the argument-scaled tolerance."""
import os

import numpy as np
import pytest
import torch

from util import GOLDEN
from oracle import oracle as O

CASES = [(3, 4), (3, 6), (2, 10), (1, 1)]


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "freq.npz"))


def arg_scale(x, D, deg):
    """|argument| of every output column (1 for the pass-through block)."""
    cols = [np.ones_like(x)]
    for f in range(deg):
        cols += [np.abs(x) * 2.0 ** f] * 2
    return np.concatenate(cols, -1)


@pytest.mark.parametrize("D,deg", CASES)
def test_oracle_matches_reference(g, D, deg):
    x, y, gr, gx = g[f"x_{D}_{deg}"], g[f"y_{D}_{deg}"], g[f"g_{D}_{deg}"], g[f"gx_{D}_{deg}"]
    oy = O.freq_encode(x, deg)
    assert oy.shape == y.shape == (257, D + 2 * deg * D)
    assert np.abs(oy - y).max() <= 1e-6
    assert np.abs(O.freq_encode_backward(gr, y, D, deg) - gx).max() <= 1e-4 * max(1.0, np.abs(gx).max())


def test_module_contract():
    from freqencoder import FreqEncoder
    enc = FreqEncoder(input_dim=3, degree=6)
    assert enc.output_dim == 39 and list(enc.parameters()) == [] and "degree=6" in repr(enc)
    with pytest.raises(RuntimeError):
        enc(torch.zeros(4, 3))          # CPU tensor: no CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("D,deg", CASES)
def test_kernel_vs_oracle_and_reference(g, D, deg):
    from freqencoder import FreqEncoder
    x, y, gr, gx = g[f"x_{D}_{deg}"], g[f"y_{D}_{deg}"], g[f"g_{D}_{deg}"], g[f"gx_{D}_{deg}"]
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    enc = FreqEncoder(input_dim=D, degree=deg)
    out = enc(xt.view(1, -1, D))                       # arbitrary prefix shape
    assert tuple(out.shape) == (1, 257, D + 2 * deg * D)
    out.backward(torch.from_numpy(gr).cuda().view(1, 257, -1))
    tol = 2e-6 * (1 + arg_scale(x, D, deg))
    assert (np.abs(out[0].detach().cpu().numpy() - y) <= tol).all()
    assert (np.abs(out[0].detach().cpu().numpy() - O.freq_encode(x, deg)) <= tol).all()
    gtol = 4e-6 * (4.0 ** deg) * max(1.0, np.abs(gr).max()) + 1e-5
    assert np.abs(xt.grad.cpu().numpy() - gx).max() <= gtol
    # backward is exact given the forward outputs it reads
    ob = O.freq_encode_backward(gr, out[0].detach().cpu().numpy(), D, deg)
    assert np.abs(xt.grad.cpu().numpy() - ob).max() <= 1e-5 * max(1.0, np.abs(ob).max())


@pytest.mark.gpu
def test_kernel_sizes_and_errors():
    import _ngp_b200 as nb
    from freqencoder import freq_encode
    assert tuple(freq_encode(torch.zeros(0, 3, device="cuda"), 4, 27).shape) == (0, 27)
    big = torch.rand(1 << 20, 3, device="cuda")
    y = freq_encode(big, 4, 27)
    assert torch.equal(y[:, :3], big) and float((y[:, 3:6] - torch.sin(big)).abs().max()) < 5e-6
    with pytest.raises(RuntimeError, match="output_dim"):
        nb.call("ngp_freq_encode_forward", big.data_ptr(), 16, 3, 4, 26, y.data_ptr())
    with pytest.raises(RuntimeError, match="input_dim"):
        nb.call("ngp_freq_encode_forward", big.data_ptr(), 16, 9, 1, 27, y.data_ptr())
