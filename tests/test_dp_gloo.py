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


def _occ_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import types
    import ngp_dp
    torch.manual_seed(100 + rank)                      # diverged per-rank RNG, as in training
    m = types.SimpleNamespace(density_grid=torch.rand(2, 64), density_bitfield=(torch.rand(16) * 255).to(torch.uint8),
                              mean_density=float(rank) + 0.25)
    ngp_dp.sync_occupancy(m, src=0)
    st = ngp_dp.seed_lock(step=48)
    locked = torch.rand(4)
    ngp_dp.seed_unlock(st)
    after = torch.rand(4)                               # the per-rank stream continues where it was
    q.put((rank, m.density_grid.clone(), m.density_bitfield.clone(), m.mean_density, locked, after))
    dist.barrier()
    dist.destroy_process_group()


def test_occupancy_sync_and_seed_lock():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 17) % 1000
    procs = [ctx.Process(target=_occ_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, g0, b0, md0, l0, a0), (_, g1, b1, md1, l1, a1) = res
    assert torch.equal(g0, g1) and torch.equal(b0, b1) and md0 == md1 == 0.25
    assert torch.equal(l0, l1)                          # locked draws identical on both ranks
    assert not torch.equal(a0, a1)                      # per-rank streams restored



def test_exchange_shards_partition_the_flat_parameter_space():
    """Host logic of the fused exchange (ngp_dp.shard_bounds / segment_pieces): shards tile [0, n) on multiples of 8 and their
    pieces tile every parameter segment exactly once."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-ngp_b200")]
    import ngp_dp
    segs = [(0, 12239728), (12239728, 7168), (12246896, 11264)]
    n = 12258160
    for world in (1, 2, 3, 4, 7, 8, 16):
        b = ngp_dp.shard_bounds(n, world)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1
        assert all(x % 8 == 0 for x in b) and all(b[i] <= b[i + 1] for i in range(world))
        sizes = [b[i + 1] - b[i] for i in range(world)]
        assert max(sizes) - min(sizes) <= 8
        covered = [0] * len(segs)
        for r in range(world):
            for i, a, cnt in ngp_dp.segment_pieces(segs, b[r], b[r + 1]):
                off, k = segs[i]
                assert off <= a and a + cnt <= off + k and a % 8 == 0 and cnt % 8 == 0
                covered[i] += cnt
        assert covered == [k for _, k in segs]
    import pytest
    with pytest.raises(ValueError):
        ngp_dp.shard_bounds(1001, 2)
