"""ngp_dp.py — ray-sharded data parallelism for the hot path (new functionality, the reference has
no live distributed path).

One process per GPU.  Every rank holds a full replica of the field (hash table 49 MB fp32 + two MLPs); a step's
rays are split contiguously across ranks; there is exactly ONE exchange per step: a sum-allreduce over a single flat
fp32 bucket {d embeddings (12 239 728), d sigma_net.weights (7168), d color_net.weights (11 264)}.  Parameters'
.grad tensors are *views into the bucket*, so autograd accumulates straight into the communication buffer (no
gather/scatter copies) and the optimizer reads the reduced values in place.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous ray range of `rank`: [lo, hi) with sizes differing by at most 1."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_indices(n, rank, world, block=256):
    """Round-robin blocks of `block` rays: rank r owns blocks r, r+world, ...  Balances per-rank sample counts (a
    contiguous split of an image gives one rank the sky and another the object)."""
    idx = torch.arange(n)
    blk = idx // block
    return idx[blk % world == rank]


class FlatGradBucket:
    def __init__(self, params, dtype=torch.float32):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=dtype, device=dev)
        off = 0
        self.views = []
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(v)
            off += p.numel()
        self.attach()

    def attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        self.flat.zero_()
        self.attach()

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def allreduce(self, group=None, average=True, async_op=False):
        """Single collective per step.  NCCL on GPUs (NVLS in-switch reduction when available), gloo in CPU tests."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average and not async_op:
            self.flat.div_(dist.get_world_size(group))
        return work


def broadcast_module(module, src=0, group=None):
    """Make every replica identical (parameters and buffers such as the occupancy bitfield)."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
