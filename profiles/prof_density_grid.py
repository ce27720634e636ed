"""Time the occupancy-grid maintenance on one GPU: fused kernels (density_grid.update_extra_state) vs the
reference's torch-op sequence (NeRFFieldFF.update_extra_state_unfused, density through GridEncoder -> FFMLP -> trunc_exp).
Usage: python profiles/prof_density_grid.py > gpurun_out/density_grid.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "torch-ngp_b200")]
import _ngp_b200 as nb                 # noqa: E402
import ngp_synth as S                  # noqa: E402
from nerf_step import NeRFFieldFF      # noqa: E402


def timed(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = {}
    for bound in (1, 2):
        torch.manual_seed(0)
        m = NeRFFieldFF(bound=bound, fused=True).cuda()
        with torch.no_grad():
            m.encoder.embeddings.uniform_(-1, 1)
        poses, intr = S.make_cameras(100, seed=11), S.intrinsics()
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            m.mark_untrained_grid(poses, intr)
            t_mark = timed(lambda: m.mark_untrained_grid(poses, intr), 5)
        res = {"cascades": m.cascade, "cells": m.cascade * 128 ** 3, "mark_untrained_ms": t_mark}
        for mode, it0 in (("full", 0), ("partial", 16)):
            def fused():
                m.iter_density = it0
                m.update_extra_state()

            def unfused():
                m.iter_density = it0
                with torch.autocast("cuda", dtype=torch.float16):
                    m.update_extra_state_unfused()
            m.iter_density = 0
            m.update_extra_state()              # make sure occupied cells exist
            for f in (fused, unfused):
                f(); f()
            nb.reset_launch_count()
            fused()
            launches = nb.launch_count()
            res[mode] = {"fused_ms": timed(fused, 10), "unfused_torch_ms": timed(unfused, 5), "fused_launches": launches}
        out[f"bound{bound}"] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
