"""GPU parity: hash-grid encoder (csrc/grid.cu) vs the CPU oracle and vs the reference's own CUDA extension."""
import numpy as np
import pytest
import torch

from util import gen, rel_err

pytestmark = pytest.mark.gpu


def _setup(B, C=2, L=16, log2T=19, desired=2048, seed=0, std=1.0, oob_frac=0.0):
    from oracle import oracle as O
    offsets, pls = O.grid_offsets(3, L, C, 2, 16, log2T, desired)
    x = torch.rand(B, 3, generator=gen(seed))
    if oob_frac > 0:
        n = int(B * oob_frac)
        x[:n] = x[:n] * 3 - 1
    table = (torch.rand(int(offsets[-1]), C, generator=gen(seed + 1)) * 2 - 1) * std
    return offsets, float(pls), x, table


def _device_scales(L, S, H):
    import _ngp_b200 as nb
    out = torch.empty(L, device="cuda")
    nb.call("ngp_grid_level_scales", out.data_ptr(), L, float(S), H)
    return out.cpu().numpy()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("B", [1, 77, 4096])
def test_forward_vs_oracle(dtype, B):
    """Indices are integer work -> bit-exact; with the device's own level scales the fp32/fp16 features are
    bit-exact too (same fma / rounding sequence)."""
    from oracle import oracle as O
    from gridencoder.grid import grid_encode
    offsets, pls, x, table = _setup(B, oob_frac=0.1 if B > 1 else 0)
    S = float(np.log2(pls))
    scales = _device_scales(16, S, 16)
    t = table.to(dtype)
    ref = O.grid_forward(x.numpy(), t.numpy(), offsets, S, 16, scales=scales)
    emb = t.cuda()
    out = grid_encode(x.cuda(), emb, torch.from_numpy(offsets).cuda(), pls, 16, False, 0, False, 0)
    assert out.dtype == dtype and out.shape == (B, 32)
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint16 if dtype == torch.float16 else np.uint32),
                                  ref.view(np.uint16 if dtype == torch.float16 else np.uint32))


def test_level_scales_close_to_libm():
    """The only non-reproducible piece (MUFU.EX2) stays within 2 ulp of libm's exp2f."""
    from oracle import oracle as O
    S = float(np.log2(np.exp2(np.log2(2048 / 16) / 15)))
    dev = _device_scales(16, S, 16)
    cpu = O.grid_level_scales(16, S, 16)
    assert np.abs(dev - cpu).max() <= 4 * np.spacing(np.float32(2048))


@pytest.mark.parametrize("cfg", [dict(C=2, L=16), dict(C=4, L=8), dict(C=8, L=4), dict(C=1, L=12), dict(C=2, L=3)])
def test_forward_shapes(cfg):
    from oracle import oracle as O
    from gridencoder.grid import grid_encode
    C, L = cfg["C"], cfg["L"]
    offsets, pls, x, table = _setup(1000, C=C, L=L, log2T=14, desired=256)
    S = float(np.log2(pls))
    scales = _device_scales(L, S, 16)
    ref = O.grid_forward(x.numpy(), table.numpy(), offsets, S, 16, scales=scales)
    out = grid_encode(x.cuda(), table.cuda(), torch.from_numpy(offsets).cuda(), pls, 16, False, 0, False, 0)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    # tiled grid + align_corners + smoothstep variants (tolerance: smoothstep polynomial contraction may differ)
    for gridtype, ac, interp in [(1, False, 0), (0, True, 0), (0, False, 1)]:
        ref = O.grid_forward(x.numpy(), table.numpy(), offsets, S, 16, gridtype=gridtype, align_corners=ac, interp=interp, scales=scales)
        out = grid_encode(x.cuda(), table.cuda(), torch.from_numpy(offsets).cuda(), pls, 16, False, gridtype, ac, interp)
        assert rel_err(out.cpu().numpy(), ref) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_backward_vs_oracle(dtype):
    """Scatter-add vs the order-independent float64 oracle sum.  fp32 atomics: 1e-5 rel; fp16 atomics (the
    reference's autocast path): bounded by fp16 accumulation, 2e-2 of the max entry at B=8192."""
    from oracle import oracle as O
    import _ngp_b200 as nb
    B = 8192
    offsets, pls, x, table = _setup(B, oob_frac=0.05)
    S = float(np.log2(pls))
    scales = _device_scales(16, S, 16)
    g = torch.randn(B, 32, generator=gen(2)).to(dtype)
    ref = O.grid_backward(g.numpy(), x.numpy(), offsets, int(offsets[-1]), 2, S, 16, scales=scales)
    ge = torch.zeros(int(offsets[-1]), 2, dtype=dtype, device="cuda")
    xd, gd, od = x.cuda(), g.cuda(), torch.from_numpy(offsets).cuda()
    nb.call("ngp_grid_encode_backward", gd.data_ptr(), xd.data_ptr(), None, od.data_ptr(), ge.data_ptr(), B, 3, 2, 16,
            S, 16, None, None, 0, 0, 0, 1 if dtype == torch.float16 else 0, 0)
    got = ge.float().cpu().numpy()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel_err(got, ref) < tol
    # no stray writes: entries the oracle never touches stay exactly zero (index parity of the scatter)
    assert not np.any(got[ref == 0] != 0)


def test_autograd_module_and_input_grad():
    """GridEncoder module: autocast fp16 path, gradient w.r.t. embeddings and inputs (dy_dx path)."""
    from oracle import oracle as O
    from gridencoder import GridEncoder
    torch.manual_seed(0)
    enc = GridEncoder(desired_resolution=2048).cuda()
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    x = (torch.rand(2048, 3, generator=gen(5)) * 2 - 1).cuda()
    S = float(np.log2(enc.per_level_scale))
    scales = _device_scales(16, S, 16)
    # fp32, inputs require grad
    xr = x.clone().requires_grad_(True)
    y = enc(xr)
    assert y.dtype == torch.float32 and y.shape == (2048, 32)
    g = torch.randn(2048, 32, generator=gen(6)).cuda()
    y.backward(g)
    x01 = ((x + 1) / 2).cpu().numpy()
    ref, dy = O.grid_forward(x01, enc.embeddings.detach().cpu().numpy(), enc.offsets.cpu().numpy(), S, 16,
                             scales=scales, want_dy_dx=True)
    np.testing.assert_array_equal(y.detach().cpu().numpy(), ref)
    gin_ref = (g.cpu().numpy().reshape(2048, 16, 1, 2) * dy.reshape(2048, 16, 3, 2)).sum(axis=(1, 3)) * 0.5
    assert rel_err(xr.grad.cpu().numpy(), gin_ref) < 1e-4
    gref = O.grid_backward(g.cpu().numpy(), x01, enc.offsets.cpu().numpy(), enc.embeddings.shape[0], 2, S, 16, scales=scales)
    assert rel_err(enc.embeddings.grad.cpu().numpy(), gref) < 1e-5
    # autocast: half output, fp32 .grad on the parameter
    enc.zero_grad()
    with torch.autocast("cuda", dtype=torch.float16):
        yh = enc(x)
    assert yh.dtype == torch.float16
    refh = O.grid_forward(x01, enc.embeddings.detach().half().cpu().numpy(), enc.offsets.cpu().numpy(), S, 16, scales=scales)
    np.testing.assert_array_equal(yh.detach().cpu().numpy().view(np.uint16), refh.view(np.uint16))
    yh.backward(g.half())
    assert enc.embeddings.grad.dtype == torch.float32


def test_vs_reference_extension():
    """Same inputs through the reference's own kernels (oracle/_ref): features bit-identical (fp32 & fp16),
    fp32 grads within atomics-order noise."""
    from oracle import ref_driver as R
    if not R.available("gridencoder"):
        pytest.skip("oracle/_ref/gridencoder not built")
    from gridencoder.grid import grid_encode
    import _ngp_b200 as nb
    B = 65536
    offsets, pls, x, table = _setup(B, oob_frac=0.02)
    od = torch.from_numpy(offsets).cuda()
    xd = x.cuda()
    for dtype in (torch.float32, torch.float16):
        emb = table.to(dtype).cuda()
        ref, _ = R.grid_encode_forward(xd, emb, od, pls, 16)
        out = grid_encode(xd, emb, od, pls, 16, False, 0, False, 0)
        assert torch.equal(out, ref), f"forward mismatch for {dtype}"
        g = torch.randn(B, 32, generator=gen(2)).to(dtype).cuda()
        gref, _ = R.grid_encode_backward(g, xd, emb, od, pls, 16)
        ge = torch.zeros_like(emb)
        nb.call("ngp_grid_encode_backward", g.data_ptr(), xd.data_ptr(), None, od.data_ptr(), ge.data_ptr(), B, 3, 2, 16,
                float(np.log2(pls)), 16, None, None, 0, 0, 0, 1 if dtype == torch.float16 else 0, 0)
        tol = 1e-5 if dtype == torch.float32 else 5e-2
        assert rel_err(ge.float().cpu().numpy(), gref.float().cpu().numpy()) < tol


def test_full_size_properties():
    """BASELINE config sizes (T=2^19, 1M points): linearity in the table and adjointness <enc(x;T), g> = <T, bwd(g)>."""
    from gridencoder.grid import grid_encode
    import _ngp_b200 as nb
    B = 1 << 20
    offsets, pls, x, table = _setup(B)
    od = torch.from_numpy(offsets).cuda(); xd = x.cuda()
    t1 = table.cuda(); t2 = torch.randn_like(t1)
    f = lambda t: grid_encode(xd, t, od, pls, 16, False, 0, False, 0)
    y1, y2, y12 = f(t1), f(t2), f(t1 + 2 * t2)
    assert rel_err((y1 + 2 * y2).cpu().numpy(), y12.cpu().numpy()) < 1e-5
    g = torch.randn(B, 32, device="cuda")
    ge = torch.zeros_like(t1)
    nb.call("ngp_grid_encode_backward", g.data_ptr(), xd.data_ptr(), None, od.data_ptr(), ge.data_ptr(), B, 3, 2, 16,
            float(np.log2(pls)), 16, None, None, 0, 0, 0, 0, 0)
    lhs = (y1.double() * g.double()).sum().item()
    rhs = (t1.double() * ge.double()).sum().item()
    assert abs(lhs - rhs) / abs(lhs) < 1e-4


def test_empty_and_errors():
    from gridencoder.grid import grid_encode
    offsets, pls, x, table = _setup(4)
    od = torch.from_numpy(offsets).cuda()
    out = grid_encode(torch.zeros(0, 3, device="cuda"), table.cuda(), od, pls, 16, False, 0, False, 0)
    assert out.shape == (0, 32)
    with pytest.raises(RuntimeError):
        grid_encode(x.cuda(), torch.rand(int(offsets[-1]), 3, device="cuda"), od, pls, 16, False, 0, False, 0)  # C=3 unsupported
