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



# ------------------------------------------------------------------------------------------------ peer-memory exchange
def shard_bounds(n, world, align=8):
    """Flat-parameter shards for the fused exchange + optimizer (csrc/exchange.cu): `world` contiguous ranges covering [0, n) whose
    boundaries are multiples of `align` elements (n itself must be).  Returns [lo_0, lo_1, ..., lo_world = n]."""
    if n % align:
        raise ValueError(f"shard_bounds: n = {n} is not a multiple of {align}")
    units = n // align
    base, rem = divmod(units, world)
    out = [0]
    for r in range(world):
        out.append(out[-1] + (base + (1 if r < rem else 0)) * align)
    return out


def segment_pieces(segments, lo, hi):
    """Intersection of the flat range [lo, hi) with parameter segments [(off, count), ...]: [(segment index, piece_lo, piece_count)]."""
    out = []
    for i, (off, cnt) in enumerate(segments):
        a, b = max(lo, off), min(hi, off + cnt)
        if b > a:
            out.append((i, a, b - a))
    return out


class _RawCuda:
    """Device memory owned by libngp_b200 (cudaMalloc / CUDA IPC mapping), exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerExchange:
    """Peer-visible buffers of every rank of one node, for the fused gradient exchange + optimizer (csrc/exchange.cu).

    Each rank allocates ONE block through the library (ngp_peer_alloc: cudaMalloc + CUDA IPC handle) holding
        [ signal pad 16 KB | fp16 gradient sink (n) | fp16 operand shadow (n) ]
    the 64-byte handles travel through torch.distributed (all_gather_object — plumbing), and every rank maps the other ranks' blocks
    (ngp_peer_open, peer access over NVLink).  `sink` / `shadow` are torch views of the LOCAL block; `pads` / `sinks` / `shadows` are
    host arrays of the per-rank device pointers as mapped in this process, handed to the exchange kernels by value."""

    def __init__(self, n_elems, group=None, device=None):
        import ctypes
        import _ngp_b200 as nb
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n = int(n_elems)
        if self.n % 8:
            raise ValueError("PeerExchange: element count must be a multiple of 8")
        if self.world > 16:
            raise RuntimeError("PeerExchange: at most 16 ranks per node")
        lib = nb.load()
        self._lib = lib
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        pad = int(lib.ngp_exchange_pad_bytes())
        seg = (2 * self.n + 255) // 256 * 256
        self._pad_bytes, self._seg_bytes = pad, seg
        total = pad + 2 * seg
        self._symm = None
        self.mc_sink = self.mc_shadow = None
        self.memory = None
        import os
        mode = os.environ.get("NGP_EXCHANGE_MEM", "auto")        # auto | symm | ipc
        if mode in ("auto", "symm"):
            # torch's symmetric-memory allocator (plumbing): cuMem allocations exchanged between the ranks AND bound to one NVSwitch
            # multicast object, which is what the NVLS forms of the exchange kernels need (multimem.ld_reduce / multimem.st)
            try:
                self._init_symm(total, pad, seg, group, dev)
            except Exception as e:
                if mode == "symm":
                    raise
                self._symm = None
                self._symm_error = f"{type(e).__name__}: {e}"
        # every rank must end up on the same kind of memory
        ok = torch.tensor([1 if self._symm is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self._symm = None
            self.mc_sink = self.mc_shadow = None
            self._init_ipc(total, pad, seg, group, dev)
        arr = ctypes.c_void_p * self.world
        self.pads = arr(*[b for b in self._base])
        self.sinks = arr(*[b + pad for b in self._base])
        self.shadows = arr(*[b + pad + seg for b in self._base])
        assert self.sink.data_ptr() == self._base[self.rank] + pad and self.sink.dtype == torch.half
        self.bounds = shard_bounds(self.n, self.world)
        dist.barrier(group=group)        # every rank has mapped every block before anyone signals into it

    def _init_symm(self, total, pad, seg, group, dev):
        import torch.distributed._symmetric_memory as symm
        grp = group if group is not None else dist.group.WORLD
        buf = symm.empty(total, dtype=torch.uint8, device=dev)
        hdl = symm.rendezvous(buf, grp)
        buf.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        self._base = [int(p) for p in hdl.buffer_ptrs]
        assert self._base[self.rank] == buf.data_ptr()
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if mc:
            self.mc_sink, self.mc_shadow = mc + pad, mc + pad + seg
        self.sink = buf[pad:pad + 2 * self.n].view(torch.half)
        self.shadow = buf[pad + seg:pad + seg + 2 * self.n].view(torch.half)
        self._symm = (buf, hdl)
        self.memory = "torch symmetric memory" + (" + NVSwitch multicast (NVLS)" if mc else " (no multicast)")

    def _init_ipc(self, total, pad, seg, group, dev):
        import ctypes
        lib = self._lib
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(dev):
            rc = lib.ngp_peer_alloc(total, ctypes.byref(ptr), handle)
            if rc != 0:
                raise RuntimeError(f"ngp_peer_alloc failed ({rc}): {lib.ngp_last_error().decode()}")
            handles = [None] * self.world
            dist.all_gather_object(handles, (bytes(handle.raw), int(dev.index)), group=group)
            self._base = [None] * self.world
            self._base[self.rank] = int(ptr.value)
            for r, (h, _) in enumerate(handles):
                if r == self.rank:
                    continue
                p = ctypes.c_void_p()
                rc = lib.ngp_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p))
                if rc != 0:
                    raise RuntimeError(f"ngp_peer_open(rank {r}) failed ({rc}): {lib.ngp_last_error().decode()}")
                self._base[r] = int(p.value)
        self._raw = (_RawCuda(self._base[self.rank] + pad, self.n, "<f2"), _RawCuda(self._base[self.rank] + pad + seg, self.n, "<f2"))
        self.sink = torch.as_tensor(self._raw[0], device=dev)
        self.shadow = torch.as_tensor(self._raw[1], device=dev)
        self.memory = "cudaMalloc + CUDA IPC (library-owned)"

    @property
    def my_range(self):
        return self.bounds[self.rank], self.bounds[self.rank + 1]

    def barrier(self, slot, flag_in=None, flag_out=None, timeout_ms=20000):
        import _ngp_b200 as nb
        nb.call("ngp_exchange_barrier", self.pads, self.rank, self.world, int(slot), flag_in, flag_out, int(timeout_ms))

    def error(self):
        """Non-zero when a barrier timed out (1 + slot).  Host read: synchronises."""
        import ctypes
        torch.cuda.synchronize(self.device)
        out = ctypes.c_uint32(0)
        rc = self._lib.ngp_exchange_error(ctypes.c_void_p(self._base[self.rank]), ctypes.byref(out))
        if rc != 0:
            raise RuntimeError(f"ngp_exchange_error failed: {self._lib.ngp_last_error().decode()}")
        return int(out.value)

    def close(self):
        """Unmap the peers' blocks and free the local one (after a barrier: nobody may still be signalling into it)."""
        if getattr(self, "_base", None) is None:
            return
        import ctypes
        torch.cuda.synchronize(self.device)
        try:
            dist.barrier(group=self.group)
        except Exception:
            pass
        if self._symm is not None:
            self.sink = self.shadow = None
            self._symm = None
            self._base = None
            return
        for r, b in enumerate(self._base):
            if r != self.rank and b:
                self._lib.ngp_peer_close(ctypes.c_void_p(b))
        self.sink = self.shadow = None
        self._lib.ngp_peer_free(ctypes.c_void_p(self._base[self.rank]))
        self._base = None
