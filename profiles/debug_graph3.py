import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import _ngp_b200 as nb
dev = "cuda"
N = 4096
ro = torch.randn(N, 3, device=dev); rd = torch.randn(N, 3, device=dev); aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0], device=dev)
nears = torch.empty(N, device=dev); fars = torch.empty(N, device=dev)

class F(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        return a * 2
    @staticmethod
    def backward(ctx, g):
        nb.call("ngp_near_far_from_aabb", ro.data_ptr(), rd.data_ptr(), aabb.data_ptr(), N, 0.2, nears.data_ptr(), fars.data_ptr())
        return g * 2

class G(torch.autograd.Function):      # same but allocating + memset-like torch op in backward
    @staticmethod
    def forward(ctx, a):
        return a * 2
    @staticmethod
    def backward(ctx, g):
        t = torch.zeros(1000, device=dev)
        return g * 2 + t.sum()

a = torch.randn(100, device=dev, requires_grad=True)
def f_ctypes_in_bwd(): F.apply(a).sum().backward()
def f_alloc_in_bwd(): G.apply(a).sum().backward()

import bench
from ngp_optim import FusedFieldOptimizer
model, _ = bench.build_model(torch.device("cuda", 0), fused=True)
xyzs = (torch.rand(128 * 100, 3, device=dev) * 2 - 1); dirs = torch.randn(128 * 100, 3, device=dev); dirs = dirs / dirs.norm(dim=-1, keepdim=True)
def f_field_nosink():
    with torch.autocast("cuda", dtype=torch.float16):
        s, c = model(xyzs, dirs)
    (s.sum() * 1e-3 + c.sum()).backward()

xb = (torch.rand(530001, 3, device=dev) * 2 - 1); db = torch.randn(530001, 3, device=dev); db = db / db.norm(dim=-1, keepdim=True)
def f_field_big():
    with torch.autocast("cuda", dtype=torch.float16):
        s, c = model(xb, db)
    (s.sum() * 1e-3 + c.sum()).backward()
fopt = None
def install():
    global fopt
    fopt = FusedFieldOptimizer(model.encoder, model.sigma_net, model.color_net, init_scale=128.0)
def f_field_sink():
    with torch.autocast("cuda", dtype=torch.float16):
        s, c = model(xyzs, dirs)
    (s.sum() * 1e-3 + c.sum()).backward()
def f_field_sink_opt():
    f_field_sink(); fopt.step()
tests = [("field_nosink_small", f_field_nosink), ("field_nosink_big_ragged", f_field_big), ("INSTALL", install), ("field_sink_small", f_field_sink), ("field_sink_opt", f_field_sink_opt), ("field_sink_big", f_field_big)]
for name, fn in tests:
    if name == "INSTALL":
        fn(); continue
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
        print(f"{name}: FAILED {type(e).__name__}: {str(e)[:300]}", flush=True)
        try: torch.cuda.synchronize()
        except Exception as e2: print("  sync:", str(e2)[:100])
