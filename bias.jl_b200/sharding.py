"""Multi-GPU data parallelism for the assignment sweep (AD-LDA): groups (documents) are split
into contiguous shards, one per rank; every rank keeps a full replica of the shared component
statistics (n_wk, n_k / sum x, n) and samples its shard against it.  Once per sweep the ranks
all-reduce the count DELTAS (NCCL over NVLink on GPUs; any torch.distributed backend works,
the CPU tests use gloo) and rebase.  The reference is single-process; this
is the one exchange step the path needs, there is no other collective.
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
        typestr = {"int32": "<i4", "float64": "<f8"}[dtype]
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, n: int, dtype: str, device):
    import torch
    return torch.as_tensor(_CudaView(ptr, n, dtype), device=device)


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
        self.merge_local_counts()

    def merge_local_counts(self):
        """After (re)building the counts from the local shard only: sum them into the global replica."""
        if self.world > 1:
            self.shard.delta_begin_zero()
            self._exchange()

    def _exchange(self):
        for t in self.shard.delta_tensors():
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        self.shard.delta_apply()

    def sweep(self, n_sweeps: int = 1):
        for _ in range(n_sweeps):
            self.shard.sweep(1)
            if self.world > 1:
                self._exchange()        # delta_apply leaves base == live, ready for the next sweep


class GpuShard:
    """Adapter: ResidentSampler -> the ShardedSampler protocol, with the delta buffers exposed as
    torch tensors so torch.distributed (NCCL) can reduce them in place on the sampler's stream."""

    def __init__(self, resident, device):
        self.rs, self.device = resident, device

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
