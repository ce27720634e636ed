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
@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_two_rank_reduced_sink_equals_single_gpu_sink(tmp_path, exchange):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dp_sink_worker import sink_after_backward
    out = str(tmp_path / "sink2.pt")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541" if exchange == "nccl" else "29547", os.path.join(ROOT, "tests", "dp_sink_worker.py"), out, exchange],
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


def _run_exchange_worker(mode, out, port):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "dp_exchange_worker.py"), out, mode],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-1500:], p.stderr[-3000:])
    return torch.load(out)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under `gpurun --gpus 2`)")
def test_peer_exchange_matches_nccl(tmp_path):
    """The fused peer-memory exchange + sharded optimizer (csrc/exchange.cu) against NCCL all-reduce + the full optimizer pass on the same
    gradients: 3 steps, the second one non-finite on one rank (skipped everywhere, loss scale halved)."""
    peer = _run_exchange_worker("peer", str(tmp_path / "peer.pt"), 29543)
    nccl = _run_exchange_worker("nccl", str(tmp_path / "nccl.pt"), 29545)
    assert peer["scale"] == nccl["scale"] == 512.0                 # one backoff
    assert int(peer["state"][3]) == int(nccl["state"][3]) == 2      # two steps actually taken
    assert torch.equal(peer["shadow0"], peer["shadow1"])            # the skipped step changed nothing
    # peer loads: two summands, fp32-accumulated and fp16 sums round identically and the Adam arithmetic is the same code -> identical.
    # NVLS (multimem.ld_reduce): the sum is formed inside the NVSwitch, whose rounding of an fp16 pair sum is not RN(fp32 sum) for every
    # input (measured here: differences of one fp16 ulp of the summed gradient on a small fraction of the entries) -> 1-ulp tolerance.
    tol = 5e-3 if peer["nvls"] else 1e-6
    for k in ("params", "exp_avg", "exp_avg_sq"):
        a, b = peer[k].double(), nccl[k].double()
        err = float((a - b).abs().max() / b.abs().max())
        print(f"peer ({peer['memory']}) vs nccl: {k} max err {err:.3e} of max")
        assert err < tol, (k, err)
    same = float((peer["shadow"] == nccl["shadow"]).float().mean())
    assert same > (0.99 if peer["nvls"] else 0.99999), same
    print(f"peer vs nccl exchange: identical fp16 operand entries {same:.6f}")
