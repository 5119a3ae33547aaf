"""Multi-GPU data parallelism for the assignment sweep (AD-LDA): groups (documents) are split
into contiguous shards, one per rank; every rank keeps a full replica of the shared component
statistics (n_wk, n_k / sum x, n) and samples its shard against it, and once per sweep the
replicas are reconciled.  The reference is single-process; this is the one
exchange step the path needs, there is no other collective.

* LDA x word ids (the headline): the exchange lives INSIDE the library (csrc/comm.cu: the library's
  own reduce-scatter / all-gather kernel over NVLink peer memory).  `attach_ranks` below makes the
  resident model of every torchrun rank a shard of one run — torch.distributed only carries the IPC
  handles — and `models.Group` / `models.lda_run_multi` do the same for shards of one process.
* BMM and the other fixed-K models: `ShardedSampler` all-reduces the count DELTAS with
  torch.distributed (NCCL on GPUs; the CPU tests use gloo) between bias_model_delta_pack / apply.
* HDP: `ShardedHDP` (group shards, turn-taking births).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_groups: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous equal split of group indices [lo, hi) for `rank`."""
    base, rem = divmod(n_groups, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_csr(group_ptr: np.ndarray, rank: int, world: int):
    """-> (lo_group, hi_group, lo_item, hi_item, local_ptr)"""
    G = len(group_ptr) - 1
    lo, hi = shard_bounds(G, rank, world)
    i0, i1 = int(group_ptr[lo]), int(group_ptr[hi])
    return lo, hi, i0, i1, (group_ptr[lo:hi + 1] - group_ptr[lo]).astype(np.int64)


def shard_items(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """BMM: flat items split contiguously."""
    return shard_bounds(n_items, rank, world)


class _CudaView:
    """Zero-copy torch view of a raw device pointer (through __cuda_array_interface__)."""

    def __init__(self, ptr: int, n: int, dtype: str):
        typestr = {"int32": "<i4", "float64": "<f8", "int64": "<i8", "int16": "<i2"}[dtype]
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, n: int, dtype: str, device):
    import torch
    return torch.as_tensor(_CudaView(ptr, n, dtype), device=device)


class _Done:
    """Stand-in for a recorded CUDA event when the shard lives in host memory."""

    def synchronize(self):
        pass


class ShardedSampler:
    """Drives one rank's shard: delta_begin -> sweep -> delta_pack -> all_reduce -> delta_apply.

    `shard` is anything with delta_begin(), delta_begin_zero(), sweep(n), delta_tensors() -> [torch.Tensor] and
    delta_apply(): bias.jl_b200.models.ResidentSampler (through GpuShard) on B200s, or the
    oracle-backed stand-in the CPU tests use to check this logic under gloo.
    """

    def __init__(self, shard, group=None):
        import torch.distributed as dist
        self.shard, self.group, self.dist = shard, group, dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._flag_host, self._flag_event = None, None
        self.merge_local_counts()

    def merge_local_counts(self):
        """After (re)building the counts from the local shard only: sum them into the global replica."""
        self._flag_host, self._flag_event = None, None      # a flag raised before the reload says nothing about the new counts
        if self.world > 1:
            self.shard.delta_begin_zero()
            self._exchange(force32=True)        # whole counts, not deltas: they do not fit 16 bits in general

    def _exchange(self, force32: bool = False):
        # Word-topic deltas travel in 16 bits when the shard offers it (half the all-reduce bytes).  The overflow flag
        # rides in the small int32 buffer; apply16 checks it on the device and applies nothing when it is set, so the
        # host never has to wait for it: it reads the flag of exchange s when it starts exchange s+1 (a pinned-memory
        # copy finished long before), and if it was set makes that one a 32-bit exchange — the skipped deltas are still
        # in live - base and simply ride along, one sweep late.
        pack16 = getattr(self.shard, "delta_tensors16", None)
        use16 = pack16 is not None and not force32
        if use16 and self._flag_event is not None:
            self._flag_event.synchronize()      # recorded one sweep ago: returns at once
            use16 = int(self._flag_host[0]) == 0
            self._flag_event = None
        fence = getattr(self.shard, "fence", lambda before: None)
        ts = pack16(self.world) if use16 else None
        if ts is not None:
            import torch
            fence(True)
            for t in ts:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            fence(False)
            self.shard.delta_apply16()
            if ts[1].is_cuda:
                if self._flag_host is None:
                    self._flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
                self._flag_host.copy_(ts[1][-4:-3], non_blocking=True)
                self._flag_event = torch.cuda.Event()
                self._flag_event.record()
            else:                                   # host shards (the CPU tests): the flag is there already
                self._flag_host = ts[1][-4:-3].clone()
                self._flag_event = _Done()
            return
        ts = self.shard.delta_tensors()
        fence(True)
        for t in ts:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        fence(False)
        self.shard.delta_apply()

    def sweep(self, n_sweeps: int = 1):
        for _ in range(n_sweeps):
            self.shard.sweep(1)
            if self.world > 1:
                self._exchange()        # delta_apply leaves base == live, ready for the next sweep
        self.flush()

    def flush(self):
        """The replicas are consistent on return.  A 16-bit exchange whose overflow flag came up applied nothing (its
        deltas wait for the next exchange); between sweeps that costs nothing, but whoever reads the counts after the
        last sweep of a call must not see a replica that misses them: the flag of the last exchange is awaited here and,
        if it is set, the pending deltas travel in 32 bits now.  (Found by tools/bmm_sharded_check.py on two GPUs:
        1 M sentences, cells that move by more than 16,383 in one sweep.)"""
        if self.world > 1 and self._flag_event is not None:
            self._flag_event.synchronize()
            if int(self._flag_host[0]) != 0:
                self._flag_event = None
                self._exchange(force32=True)


class GpuShard:
    """Adapter: ResidentSampler -> the ShardedSampler protocol, with the delta buffers exposed as
    torch tensors so torch.distributed (NCCL) can reduce them in place on the sampler's stream."""

    def __init__(self, resident, device):
        self.rs, self.device = resident, device

    def fence(self, before: bool):
        """The library launches on its context's stream, torch.distributed on torch's current stream.  When the two
        differ (a Context created without a stream owns a private one) the pack kernels must finish before the
        collective reads the buffers, and the collective before the apply kernels read them."""
        import torch
        cur = torch.cuda.current_stream(self.device)
        if self.rs.ctx.stream_handle == cur.cuda_stream and cur.cuda_stream != 0:
            return                              # same stream: already ordered
        if before:
            self.rs.ctx.sync()
        else:
            cur.synchronize()

    def delta_begin(self):
        self.rs.delta_begin()

    def delta_begin_zero(self):
        self.rs.delta_begin_zero()

    def sweep(self, n):
        self.rs.sweep(n)

    def delta_tensors(self):
        return [device_tensor(p, n, dt, self.device) for (p, n, dt) in self.rs.delta_pack()]

    def delta_apply(self):
        self.rs.delta_apply()

    def delta_tensors16(self, world):
        packs = self.rs.delta_pack16(world)
        if packs is None or packs[0][1] == 0:
            return None
        return [device_tensor(p, n, dt, self.device) for (p, n, dt) in packs]

    def delta_apply16(self):
        self.rs.delta_apply16()


def attach_ranks(rs, n_items_cap: int, group=None):
    """One process per GPU: make the resident LDA model `rs` of every rank a shard of one run whose count exchange
    happens inside the library (csrc/comm.cu, peer memory over NVLink).  torch.distributed only carries the few bytes
    of host-side plumbing: the capacities and the IPC handles of the ranks' shared blocks.  Afterwards rs.sweep(n) is
    collective."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    caps = [None] * world
    dist.all_gather_object(caps, int(n_items_cap), group=group)
    rs.share_prepare(rank, world, int(sum(caps)))
    handles = [None] * world
    dist.all_gather_object(handles, rs.share_export(), group=group)
    rs.share_attach(handles)
    dist.barrier(group)


# ------------------------------------------------------------------ sharded HDP
class DistComm:
    """Collectives of one rank of a torch.distributed job (NCCL on GPUs)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ranks = [self.rank]          # global rank of each local shard

    def bcast_state(self, turn: int, states, device):
        """states[i] = (KK, beta) of local shard i if it is rank `turn`, else None -> (KK, beta) of rank `turn`"""
        import torch
        mine = states[0]
        if self.world == 1:
            return mine
        n = len(mine[1]) if mine is not None else 0
        nn = torch.tensor([n], dtype=torch.int64, device=device)
        self.dist.all_reduce(nn, op=self.dist.ReduceOp.MAX, group=self.group)
        buf = torch.zeros(int(nn.item()) + 1, dtype=torch.float64, device=device)
        if mine is not None:
            buf[0] = float(mine[0])
            buf[1:] = torch.from_numpy(mine[1]).to(device)
        self.dist.broadcast(buf, src=turn, group=self.group)
        h = buf.cpu().numpy()
        return int(round(h[0])), h[1:].copy()

    def allreduce(self, tensor_lists):
        if self.world == 1:
            return
        for ts in tensor_lists:           # one list per local shard (here: exactly one)
            for t in ts:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)


class LocalComm:
    """All 'ranks' are samplers of this process (tests: several shards on one GPU); sums are done in place."""

    def __init__(self, n_shards: int):
        self.world, self.ranks = n_shards, list(range(n_shards))

    def bcast_state(self, turn: int, states, device):
        return states[turn]

    def allreduce(self, tensor_lists):
        for seg in range(len(tensor_lists[0])):
            total = tensor_lists[0][seg].clone()
            for ts in tensor_lists[1:]:
                total += ts[seg]
            for ts in tensor_lists:
                ts[seg].copy_(total)


class ShardedHDP:
    """HDP direct-assignment sampler with the groups split over ranks.  Every rank keeps a full replica of
    the component statistics, beta, alpha0 and gamma; per sweep the ranks sample their groups in parallel, take turns
    giving birth to new components (all-reducing the deltas of the component statistics after every turn), and
    all-reduce the per-component table totals of each internal iteration (the protocol is spelled out in include/bias_cuda.h).  `samplers` are this process's
    ResidentSampler shards (one per process under torch.distributed; several with LocalComm)."""

    def __init__(self, samplers, comm, device, host_seed: int = 0):
        self.s, self.comm, self.device = list(samplers), comm, device
        for rs in self.s:
            rs.set_host_seed(host_seed)
        if comm.world > 1:
            # counts were built from the local groups only: merge them, then drop empty components consistently
            for rs in self.s:
                rs.delta_begin_zero()
            self._exchange_deltas()
            for rs in self.s:
                rs.hdp_shard_prune()

    def _tensors(self, packs):
        return [[device_tensor(p, n, dt, self.device) for (p, n, dt) in pk] for pk in packs]

    def _allreduce(self, packs):
        """The library works on its contexts' streams, torch / NCCL on torch's current stream: order them explicitly
        (the HDP sweep synchronises with the host at every phase anyway)."""
        import torch
        for rs in self.s:
            rs.ctx.sync()
        self.comm.allreduce(self._tensors(packs))
        torch.cuda.current_stream(self.device).synchronize()

    def _exchange_deltas(self):
        self._allreduce([rs.delta_pack() for rs in self.s])
        for rs in self.s:
            rs.delta_apply()

    def sweep(self, n_sweeps: int = 1):
        for _ in range(n_sweeps):
            for rs in self.s:
                rs.delta_begin()
            for rs in self.s:
                rs.hdp_shard_pass()
            for turn in range(self.comm.world):
                # rank `turn` gives birth against everything the ranks before it opened; the others adopt its KK / beta
                states = [rs.hdp_shard_births() if r == turn else None for rs, r in zip(self.s, self.comm.ranks)]
                KK, beta = self.comm.bcast_state(turn, states, self.device)
                for rs, r in zip(self.s, self.comm.ranks):
                    if r != turn:
                        rs.hdp_shard_adopt(KK, beta[:KK + 1])
                self._exchange_deltas()
            for rs in self.s:
                rs.hdp_shard_prune()
            for hh in range(self.s[0].n_internals):
                self._allreduce([rs.hdp_shard_tables(hh) for rs in self.s])
                for rs in self.s:
                    rs.hdp_shard_draw()
            for rs in self.s:
                rs.hdp_shard_end()
