"""Generate the committed golden vectors (tests/golden/*.npz) by running the REFERENCE's own CUDA extensions
(oracle/_ref, built from /root/reference/*/src by oracle/build_ref.py) on seeded synthetic inputs.

Run on the GPU box:  python tests/golden/make_golden.py gpurun_out/golden
then copy the .npz files into tests/golden/ and commit them.  They pin the CPU oracle (tests/test_oracle_cpu.py,
runs without a GPU) and the CUDA path (tests/test_gpu_golden.py).  Sizes are kept small (few hundred KB total).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200"), os.path.join(ROOT, "tests")]

from oracle import ref_driver as R          # noqa: E402
from oracle import oracle as O              # noqa: E402
from util import gen, synth_rays            # noqa: E402


def main(out):
    os.makedirs(out, exist_ok=True)
    dev = "cuda"
    # ---- hash grid: small table (T=2^12, L=8 incl. dense + hashed levels), 512 points, fp32 + fp16
    offsets, pls = O.grid_offsets(3, 8, 2, 2, 16, 12, 512)
    x = torch.rand(512, 3, generator=gen(0)); x[:16] = x[:16] * 3 - 1
    table = torch.rand(int(offsets[-1]), 2, generator=gen(1)) * 2 - 1
    g = torch.randn(512, 16, generator=gen(2))
    od = torch.from_numpy(offsets).to(dev)
    d = dict(offsets=offsets, per_level_scale=np.float64(pls), x=x.numpy(), table=table.numpy(), grad=g.numpy())
    for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
        emb = table.to(dt).to(dev)
        y, dy = R.grid_encode_forward(x.to(dev), emb, od, pls, 16, True)
        ge, gi = R.grid_encode_backward(g.to(dt).to(dev), x.to(dev), emb, od, pls, 16, dy)
        d[f"y_{name}"] = y.cpu().numpy(); d[f"dy_dx_{name}"] = dy.cpu().numpy()
        d[f"grad_table_{name}"] = ge.float().cpu().numpy(); d[f"grad_x_{name}"] = gi.float().cpu().numpy()
    # device level scales as the reference kernel computes them (exp2f on the GPU)
    import _ngp_b200 as nb
    sc = torch.empty(8, device=dev)
    nb.call("ngp_grid_level_scales", sc.data_ptr(), 8, float(np.log2(pls)), 16)
    d["level_scales"] = sc.cpu().numpy()
    np.savez_compressed(os.path.join(out, "grid.npz"), **d)

    # ---- SH, degree 8, 256 unit + non-unit vectors
    v = torch.randn(256, 3, generator=gen(7)); v[:128] /= v[:128].norm(dim=-1, keepdim=True)
    y, dy = R.sh_encode_forward(v.to(dev), 8, True)
    np.savez_compressed(os.path.join(out, "sh.npz"), dirs=v.numpy(), y=y.cpu().numpy(), dy_dx=dy.cpu().numpy())

    # ---- ray marching: 256 rays of the synthetic scene, bitfield stored as the list of occupied Morton cells
    N = 256
    rays_o, rays_d, bitfield, grid = synth_rays(N)
    noises = torch.rand(N, generator=gen(4))
    ro, rd, bf, nz = rays_o.to(dev), rays_d.to(dev), bitfield.to(dev), noises.to(dev)
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0], device=dev)
    nears, fars = R.near_far_from_aabb(ro, rd, aabb, 0.2)
    d = dict(rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), noises=noises.numpy(), nears=nears.cpu().numpy(),
             fars=fars.cpu().numpy(), occupied_cells=np.nonzero(grid.numpy().reshape(-1) > 0.01)[0].astype(np.int32))
    for tag, dtg in (("g0", 0.0), ("g1", 1.0 / 128)):
        M = N * 512
        xyzs, dirs, deltas, rays, counter = R.march_rays_train(ro, rd, 1.0, bf, 1, 128, nears, fars, M, nz, dtg, 1024)
        m = int(counter[0].item())
        d[f"counter_{tag}"] = counter.cpu().numpy(); d[f"rays_{tag}"] = rays.cpu().numpy()
        d[f"xyzs_{tag}"] = xyzs[:m].cpu().numpy(); d[f"deltas_{tag}"] = deltas[:m].cpu().numpy()
        if tag == "g0":
            sig = torch.rand(m, generator=gen(20)) * 30; rgb = torch.rand(m, 3, generator=gen(21))
            ws, dp, im = R.composite_rays_train_forward(sig.to(dev), rgb.to(dev), deltas[:m].contiguous(), rays, 1e-4)
            gws = torch.randn(N, generator=gen(22)); gim = torch.randn(N, 3, generator=gen(23))
            gs, gc = R.composite_rays_train_backward(gws.to(dev), gim.to(dev), sig.to(dev), rgb.to(dev), deltas[:m].contiguous(), rays, ws, im, 1e-4)
            d.update(sigmas=sig.numpy(), rgbs=rgb.numpy(), weights_sum=ws.cpu().numpy(), depth=dp.cpu().numpy(), image=im.cpu().numpy(),
                     grad_ws=gws.numpy(), grad_image=gim.numpy(), grad_sigmas=gs.cpu().numpy(), grad_rgbs=gc.cpu().numpy())
    # inference: 3 marching rounds of n_step=4
    alive = torch.arange(N, dtype=torch.int32, device=dev); rt = nears.clone()
    wsum = torch.zeros(N, device=dev); dep = torch.zeros(N, device=dev); img = torch.zeros(N, 3, device=dev)
    x, dd, l = R.march_rays(N, 4, alive, rt, ro, rd, 1.0, bf, 1, 128, nears, fars, torch.zeros(N, device=dev), 128)
    sig = torch.rand(x.shape[0], generator=gen(30)) * 40; rgb = torch.rand(x.shape[0], 3, generator=gen(31))
    R.composite_rays(N, 4, 1e-2, alive, rt, sig.to(dev), rgb.to(dev), l, wsum, dep, img)
    d.update(inf_xyzs=x.cpu().numpy(), inf_deltas=l.cpu().numpy(), inf_sigmas=sig.numpy(), inf_rgbs=rgb.numpy(),
             inf_alive=alive.cpu().numpy(), inf_t=rt.cpu().numpy(), inf_ws=wsum.cpu().numpy(), inf_depth=dep.cpu().numpy(),
             inf_image=img.cpu().numpy())
    np.savez_compressed(os.path.join(out, "raymarching.npz"), **d)

    # ---- FFMLP: sigma (32-64-64-16) and color (32-64-64-64-16) nets, 256 rows
    d = {}
    for tag, nl in (("sigma", 2), ("color", 3)):
        n = 64 * (32 + 64 * (nl - 1) + 16)
        w = ((torch.rand(n, generator=gen(50 + nl)) * 2 - 1) * np.sqrt(3 / 64)).half()
        xx = (torch.randn(256, 32, generator=gen(60 + nl)) * 0.5).half()
        gg = (torch.randn(256, 16, generator=gen(70 + nl)) * 0.05).half()
        y, fb = R.ffmlp_forward(xx.to(dev), w.to(dev), 32, 16, 64, nl)
        yi, _ = R.ffmlp_forward(xx.to(dev), w.to(dev), 32, 16, 64, nl, inference=True)
        gi, gw, bb = R.ffmlp_backward(gg.to(dev), xx.to(dev), w.to(dev), fb, 32, 16, 64, nl)
        torch.cuda.synchronize()
        d.update({f"{tag}_w": w.numpy(), f"{tag}_x": xx.numpy(), f"{tag}_g": gg.numpy(), f"{tag}_y": y.cpu().numpy(),
                  f"{tag}_y_inf": yi.cpu().numpy(), f"{tag}_fwd": fb.cpu().numpy(), f"{tag}_gi": gi.cpu().numpy(),
                  f"{tag}_gw": gw.cpu().numpy(), f"{tag}_bwd": bb.cpu().numpy()})
    np.savez_compressed(os.path.join(out, "ffmlp.npz"), **d)
    print("golden vectors written to", out, {f: os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
