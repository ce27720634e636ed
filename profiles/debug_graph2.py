"""Finer capture diagnosis: individual backward entry points."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import _ngp_b200 as nb
dev = "cuda"
M = 128 * 512
nl = 2
w = (torch.rand(64 * (32 + 64 + 16), device=dev) - 0.5).half()
x = torch.randn(M, 32, device=dev).half(); g = torch.randn(M, 16, device=dev).half()
fb = torch.empty(nl, M, 64, dtype=torch.half, device=dev); y = torch.empty(M, 16, dtype=torch.half, device=dev)
gi = torch.empty_like(x); gw = torch.empty_like(w)
nbytes = nb.load().ngp_ffmlp_backward_workspace_bytes(M, 32, 16, 64, nl)
ws = torch.empty(nbytes // 4, device=dev)
from oracle import oracle as O
offsets, pls = O.grid_offsets(3, 16, 2, 2, 16, 19, 2048)
od = torch.from_numpy(offsets).to(dev); S = float(np.log2(pls))
x01 = torch.rand(M, 3, device=dev); ge = torch.zeros(int(offsets[-1]), 2, dtype=torch.half, device=dev); gf = torch.randn(M, 32, device=dev).half()

def f_fwd(): nb.call("ngp_ffmlp_forward", x.data_ptr(), w.data_ptr(), M, 32, 16, 64, nl, 0, 6, fb.data_ptr(), y.data_ptr())
def f_bwd(): nb.call("ngp_ffmlp_backward", g.data_ptr(), x.data_ptr(), w.data_ptr(), fb.data_ptr(), M, 32, 16, 64, nl, 0, 6, 1, None, gi.data_ptr(), gw.data_ptr(), ws.data_ptr(), nbytes)
def f_gridb(): nb.call("ngp_grid_encode_backward", gf.data_ptr(), x01.data_ptr(), None, od.data_ptr(), ge.data_ptr(), M, 3, 2, 16, S, 16, None, None, 0, 0, 0, 1, 0)
def f_memset(): ws.zero_()
def f_torch_bwd():
    a = torch.randn(1000, device=dev, requires_grad=True); (a * a).sum().backward()

for name, fn in (("ffmlp_fwd", f_fwd), ("ffmlp_bwd", f_bwd), ("grid_bwd", f_gridb), ("torch_memset", f_memset), ("torch_autograd", f_torch_bwd)):
    try:
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn(); fn()
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            fn()
        torch.cuda.synchronize(); gph.replay(); torch.cuda.synchronize()
        print(f"{name}: OK", flush=True)
    except Exception as e:
        print(f"{name}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
        try: torch.cuda.synchronize()
        except Exception as e2: print("  sync:", str(e2)[:100])
