"""Which op of the step breaks CUDA-graph capture?  (diagnostic; run under gpurun)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200"), os.path.join(ROOT, "tests")]
import torch
import bench
from ngp_optim import FusedFieldOptimizer
import raymarching

dev = torch.device("cuda", 0)
model, _ = bench.build_model(dev, fused=True)
R = 65536
_, dev_in = bench.make_inputs(R, 0, 1, dev)
ro, rd, tgt = dev_in[0]
fopt = FusedFieldOptimizer(model.encoder, model.sigma_net, model.color_net, init_scale=128.0)
model.mean_count = 0
with torch.autocast("cuda", dtype=torch.float16):
    out = model.render_train(ro, rd, perturb=True)
model.mean_count = int(model.step_counter[0, 0].item()) + 1024
nears, fars = raymarching.near_far_from_aabb(ro, rd, model.aabb_train, 0.2)
xyzs, dirs, deltas, rays = raymarching.march_rays_train(ro, rd, 1, model.density_bitfield, 1, 128, nears, fars, None, model.mean_count, True, 128, False, 0, 1024)


def f_march():
    return raymarching.march_rays_train(ro, rd, 1, model.density_bitfield, 1, 128, nears, fars, model.step_counter[0], model.mean_count, True, 128, False, 0, 1024)

def f_field_fwd():
    with torch.autocast("cuda", dtype=torch.float16), torch.no_grad():
        return model(xyzs, dirs)

def f_field_fwd_bwd():
    with torch.autocast("cuda", dtype=torch.float16):
        s, c = model(xyzs, dirs)
    (s.sum() * 1e-3 + c.sum()).backward()

def f_composite():
    s = torch.rand(xyzs.shape[0], device=dev, requires_grad=True); c = torch.rand(xyzs.shape[0], 3, device=dev, requires_grad=True)
    w, d, im = raymarching.composite_rays_train(s, c, deltas, rays, 1e-4)
    im.sum().backward()

def f_opt():
    fopt.step()

def f_full():
    with torch.autocast("cuda", dtype=torch.float16):
        out = model.render_train(ro, rd, perturb=True)
        loss = ((out["image"] - tgt) ** 2).sum() / (3.0 * R)
    (loss * fopt.scale_tensor()).backward()
    fopt.step()

for mode in ("global", "thread_local", "relaxed"):
    for name, fn in (("march", f_march), ("field_fwd", f_field_fwd), ("field_fwd_bwd", f_field_fwd_bwd), ("composite", f_composite), ("opt", f_opt), ("full", f_full)):
        try:
            side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn(); fn()
            torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode=mode):
                fn()
            torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
            print(f"[{mode}] {name}: OK", flush=True)
        except Exception as e:
            print(f"[{mode}] {name}: FAILED {type(e).__name__}: {str(e)[:160]}", flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("   sync after failure:", str(e2)[:100], flush=True)
