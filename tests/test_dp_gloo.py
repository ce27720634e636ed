"""CPU: the N>1 path (ray sharding + single flat-bucket allreduce) with world_size 2 over gloo."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ngp_dp
    torch.manual_seed(0)
    table = torch.nn.Parameter(torch.randn(1000, 2))
    w = torch.nn.Parameter(torch.randn(77))
    if rank == 1:
        with torch.no_grad():
            table.add_(1.0)               # deliberately diverged replica
    mod = torch.nn.ParameterList([table, w])
    ngp_dp.broadcast_module(mod)
    bucket = ngp_dp.FlatGradBucket([table, w])
    assert table.grad.data_ptr() == bucket.flat.data_ptr()
    N = 101
    lo, hi = ngp_dp.shard_range(N, rank, world)
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(0, 1000, (N,), generator=g)
    tgt = torch.randn(N, generator=g)
    bucket.zero()
    # "rays" lo:hi ; loss is a SUM over local rays so that averaging the allreduce == mean over all ranks' sums / world
    loss = ((table[idx[lo:hi]].sum(-1) * w[:1] - tgt[lo:hi]) ** 2).sum()
    loss.backward()
    assert table.grad.data_ptr() == bucket.flat.data_ptr()     # autograd accumulated in place into the bucket
    bucket.allreduce(average=False)
    q.put((rank, lo, hi, bucket.flat.clone(), table.detach().clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, lo0, hi0, f0, t0), (_, lo1, hi1, f1, t1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 51, 51, 101)
    assert torch.equal(t0, t1)                 # broadcast made replicas identical
    assert torch.allclose(f0, f1)              # both ranks hold the same reduced gradient
    # single-process reference over all rays
    torch.manual_seed(0)
    table = torch.nn.Parameter(torch.randn(1000, 2)); w = torch.nn.Parameter(torch.randn(77))
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(0, 1000, (101,), generator=g); tgt = torch.randn(101, generator=g)
    loss = ((table[idx].sum(-1) * w[:1] - tgt) ** 2).sum()
    loss.backward()
    ref = torch.cat([table.grad.reshape(-1), w.grad.reshape(-1)])
    assert torch.allclose(f0, ref, rtol=1e-5, atol=1e-5)


def test_shard_range_covers_everything():
    import ngp_dp
    for n in (0, 1, 7, 640000):
        for world in (1, 2, 3, 8):
            spans = [ngp_dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
