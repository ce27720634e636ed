"""tests/golden/freq.npz from the reference's pure-PyTorch FreqEncoder (encoding.py:5-42, imported unmodified from /root/reference,
CPU) — the formulation its CUDA freqencoder extension implements with fast intrinsics.  python tests/golden/make_golden_freq.py"""
import os
import sys

import numpy as np
import torch

sys.path.append("/root/reference")
from encoding import FreqEncoder   # noqa: E402


def main():
    out = {}
    torch.manual_seed(0)
    for D, deg in ((3, 4), (3, 6), (2, 10), (1, 1)):
        x = (torch.rand(257, D) * 2 - 1).requires_grad_(True)
        enc = FreqEncoder(input_dim=D, max_freq_log2=deg - 1, N_freqs=deg, log_sampling=True)
        y = enc(x)
        g = torch.randn_like(y)
        y.backward(g)
        out[f"x_{D}_{deg}"], out[f"y_{D}_{deg}"], out[f"g_{D}_{deg}"], out[f"gx_{D}_{deg}"] = \
            x.detach().numpy(), y.detach().numpy(), g.numpy(), x.grad.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "freq.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
