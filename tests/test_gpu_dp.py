"""GPU (>= 2 devices): the data-parallel gradient exchange.  Two ranks, each on its shard of the same 8192 rays, all-reduce their fp16
gradient sinks; the result equals the sink of one GPU processing all rays, up to the order of the fp16 sums."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under `gpurun --gpus 2`)")
def test_two_rank_reduced_sink_equals_single_gpu_sink(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dp_sink_worker import sink_after_backward
    out = str(tmp_path / "sink2.pt")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tests", "dp_sink_worker.py"), out],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    two = torch.load(out)
    one, loss1 = sink_after_backward(1, 0)
    a, b = two["sink"].double(), one.double()
    assert abs(two["loss_sum"] - loss1) < 1e-5 * max(1.0, abs(loss1))
    cos = float((a * b).sum() / (a.norm() * b.norm()))
    assert cos > 0.9995, cos
    err = float((a - b).abs().max() / b.abs().max())
    assert err < 2e-2, err                      # fp16 atomics + fp16 allreduce: sum-order differences only
    # entries no ray touched stay exactly zero on both sides
    assert bool(((a == 0) == (b == 0)).float().mean() > 0.999)
    print(f"2-rank reduced sink vs 1-GPU sink: cos={cos:.6f} max err {err:.2e} of max |g|")
