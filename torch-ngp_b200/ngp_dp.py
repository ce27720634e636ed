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


def sync_occupancy(model, src=0, group=None):
    """Keep the replicas' occupancy state identical after a rank-local update_extra_state (the update draws random
    sample positions, so replicas diverge unless the RNG streams are locked): broadcast density_grid (8 MB per cascade) and
    density_bitfield (256 KB per cascade) from `src`, plus mean_density.  Call it right after update_extra_state, i.e. every 16
    steps; alternatively call seed_lock() before the update on every rank and skip the broadcast."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.broadcast(model.density_grid, src=src, group=group)
    dist.broadcast(model.density_bitfield, src=src, group=group)
    md = torch.tensor([float(model.mean_density)], dtype=torch.float64, device=model.density_grid.device)
    dist.broadcast(md, src=src, group=group)
    model.mean_density = float(md.item())


def seed_lock(step, base_seed=0):
    """Give every rank the same RNG stream for the next update_extra_state (identical weights + identical draws = identical
    grids, no exchange needed).  Returns the previous CPU / CUDA generator states so the caller can restore its per-rank stream."""
    cpu_state = torch.get_rng_state()
    cuda_state = torch.cuda.get_rng_state() if torch.cuda.is_available() else None
    torch.manual_seed(base_seed * 1000003 + int(step))
    return cpu_state, cuda_state


def seed_unlock(states):
    cpu_state, cuda_state = states
    torch.set_rng_state(cpu_state)
    if cuda_state is not None:
        torch.cuda.set_rng_state(cuda_state)

