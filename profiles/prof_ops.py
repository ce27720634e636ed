"""prof_ops.py — stand-alone driver for ncu captures of the hot-path kernels (one launch of each op at a
steady-state-like size).  Usage (under gpurun):
  ncu --set full --clock-control none --import-source on -k regex:k_ffmlp -c 6 -o gpurun_out/prof_mlp python profiles/prof_ops.py mlp
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import _ngp_b200 as nb

what = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128 * 4096
dev = "cuda"
torch.manual_seed(0)

if what in ("mlp", "all"):
    for nl in (2, 3):
        n = 64 * (32 + 64 * (nl - 1) + 16)
        w = ((torch.rand(n, device=dev) * 2 - 1) * np.sqrt(3 / 64)).half()
        x = (torch.randn(B, 32, device=dev) * 0.5).half()
        g = (torch.randn(B, 16, device=dev) * 0.05).half()
        fb = torch.empty(nl, B, 64, dtype=torch.half, device=dev); y = torch.empty(B, 16, dtype=torch.half, device=dev)
        bb = torch.empty_like(fb); gi = torch.empty_like(x); gw = torch.empty_like(w)
        nbytes = nb.load().ngp_ffmlp_backward_workspace_bytes(B, 32, 16, 64, nl)
        ws = torch.empty(nbytes // 4, device=dev)
        for _ in range(2):
            nb.call("ngp_ffmlp_forward", x.data_ptr(), w.data_ptr(), B, 32, 16, 64, nl, 0, 6, fb.data_ptr(), y.data_ptr())
            nb.call("ngp_ffmlp_backward", g.data_ptr(), x.data_ptr(), w.data_ptr(), fb.data_ptr(), B, 32, 16, 64, nl, 0, 6, 1,
                    bb.data_ptr(), gi.data_ptr(), gw.data_ptr(), ws.data_ptr(), nbytes)
        torch.cuda.synchronize()

if what in ("grid", "all"):
    from oracle import oracle as O
    from util import synth_rays
    import raymarching
    offsets, pls = O.grid_offsets(3, 16, 2, 2, 16, 19, 2048)
    od = torch.from_numpy(offsets).to(dev)
    table = ((torch.rand(int(offsets[-1]), 2, device=dev) * 2 - 1)).half()
    # ray-coherent sample positions, as the marcher produces them
    N = 65536
    ro, rd, bf, _ = synth_rays(N)
    ro, rd, bf = ro.to(dev), rd.to(dev), bf.to(dev)
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0], device=dev)
    nears, fars = raymarching.near_far_from_aabb(ro, rd, aabb, 0.2)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(ro, rd, 1, bf, 1, 128, nears, fars, None, -1, True, 128, True, 0, 1024)
    x01 = ((xyzs + 1) / 2).contiguous()
    M = x01.shape[0]
    out = torch.empty(M, 32, dtype=torch.half, device=dev)
    g = torch.randn(M, 32, device=dev).half()
    ge = torch.zeros_like(table)
    S = float(np.log2(pls))
    for _ in range(2):
        nb.call("ngp_grid_encode_forward", x01.data_ptr(), table.data_ptr(), od.data_ptr(), out.data_ptr(), M, 3, 2, 16, S, 16, None, 0, 0, 0, 1, 0)
        nb.call("ngp_grid_encode_backward", g.data_ptr(), x01.data_ptr(), None, od.data_ptr(), ge.data_ptr(), M, 3, 2, 16, S, 16, None, None, 0, 0, 0, 1, 0)
    torch.cuda.synchronize()
    print("grid samples", M)

if what in ("field", "all"):
    # fused field forward/backward at a steady-state-like size, ray-ordered samples
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import synth_rays
    import raymarching
    from nerf_step import NeRFFieldFF
    import ngp_synth as S
    m = NeRFFieldFF(bound=1, fused=True).cuda().train()
    grid, _ = S.box_union_density(128, seed=12)
    m.density_bitfield.copy_(torch.from_numpy(S.packbits_np(grid.numpy())).cuda())
    N = 65536
    ro, rd, bf, _ = synth_rays(N)
    ro, rd = ro.to(dev), rd.to(dev)
    nears, fars = raymarching.near_far_from_aabb(ro, rd, m.aabb_train, 0.2)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(ro, rd, 1, m.density_bitfield, 1, 128, nears, fars, None, -1, True, 128, True, 0, 1024)
    for _ in range(2):
        m.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            s, c = m(xyzs, dirs)
        (s.sum() * 1e-3 + c.sum()).backward()
    torch.cuda.synchronize()
    print("field samples", xyzs.shape[0])
