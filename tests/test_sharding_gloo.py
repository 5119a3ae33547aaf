"""The N>1 path on CPU: two gloo ranks drive ShardedSampler (the same class bench.py uses over NCCL)
with an oracle-backed shard.  Checks the exchange protocol — the int32 one and the 16-bit one with its overflow flag
and one-sweep-late fallback: replicas stay identical, integer counts are conserved exactly, and the merged state
equals the counts recomputed from the gathered zz."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShard:
    """CPU stand-in for GpuShard: the shared statistics are the oracle's (mi, mm, nn) arrays."""

    def __init__(self, K, V, aa_total, alpha, ptr, words, z, seed):
        from oracle import oracle as orc
        self.state = orc.LDAState(orc.Comps.md(K, V, aa_total), orc.Data.words(words), ptr, z, alpha, with_diag=False)
        c = self.state.c
        self.live = [c.mi, c.mm, c.nn]
        self.base = [a.copy() for a in self.live]
        self.delta = [np.zeros_like(a) for a in self.live]
        self.rng = np.random.default_rng(seed)

    def delta_begin(self):
        for b_, a in zip(self.base, self.live):
            b_[...] = a

    def delta_begin_zero(self):
        for b_ in self.base:
            b_[...] = 0

    def sweep(self, n):
        for _ in range(n):
            self.state.sweep(self.rng.random(len(self.state.z)), diag_mode=0)

    def delta_tensors(self):
        for d, a, b_ in zip(self.delta, self.live, self.base):
            d[...] = a - b_
        return [torch.from_numpy(d) for d in self.delta]

    def delta_apply(self):
        for d, a, b_ in zip(self.delta, self.live, self.base):
            a[...] = b_ + d
            b_[...] = a


class OracleShard16(OracleShard):
    """The same stand-in offering the 16-bit exchange of GpuShard: two word-topic deltas per int32 as d1 * 65536 + d0
    (arithmetically, so that the all-reduced sums unpack as long as every summed delta fits 16 bits), the other deltas
    and the overflow flag (fourth from the end) in a small int32 buffer; apply16 applies nothing when the flag is set."""

    def __init__(self, *a, limit=32767, **kw):
        super().__init__(*a, **kw)
        self.limit = limit
        self.n16 = 0                     # exchanges that went through in 16 bits
        self.skipped = 0                 # 16-bit exchanges refused by the flag

    def delta_tensors16(self, world):
        d = (self.live[0] - self.base[0]).reshape(-1)
        small = np.concatenate([(self.live[1] - self.base[1]).reshape(-1), (self.live[2] - self.base[2]).reshape(-1),
                                np.zeros(4, dtype=np.int64)])
        small[-4] = int(np.abs(d).max(initial=0) * world > self.limit)
        self._packed = torch.from_numpy((d[1::2] * 65536 + d[0::2]).astype(np.int32))
        self._small = torch.from_numpy(small.astype(np.int32))
        return [self._packed, self._small]

    def delta_apply16(self):
        small = self._small.numpy().astype(np.int64)
        if small[-4] != 0:
            self.skipped += 1
            return                       # the deltas stay in live - base and ride along with the next (32-bit) exchange
        pk = self._packed.numpy().astype(np.int64)
        d0 = (pk + 32768) % 65536 - 32768
        d1 = (pk - d0) // 65536
        d = np.empty(2 * len(pk), dtype=np.int64)
        d[0::2], d[1::2] = d0, d1
        n1 = self.live[1].size
        sums = [d.reshape(self.live[0].shape), small[:n1].reshape(self.live[1].shape),
                small[n1:-4].reshape(self.live[2].shape)]
        for a, b_, sd in zip(self.live, self.base, sums):
            a[...] = b_ + sd
            b_[...] = a
        self.n16 += 1


def _worker16(rank, world, port, out_dir, limit):
    """The 16-bit exchange under gloo: with a generous limit every exchange after the initial merge is a 16-bit one;
    with a limit of 0 every 16-bit attempt is refused by the flag and the next exchange, a 32-bit one, carries the
    skipped deltas.  Either way the replicas end up equal to the counts recomputed from the gathered zz."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bias_jl_b200 as b
    from bias_jl_b200.sharding import ShardedSampler, shard_csr

    K, V = 10, 26
    xx, _, _ = b.data.bars_lda_corpus(n_groups=41, n_tokens=30, seed=3)
    words = np.concatenate(xx)
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in xx])]).astype(np.int64)
    z0 = np.random.default_rng(0).integers(0, K, len(words))
    lo, hi, i0, i1, lptr = shard_csr(ptr, rank, world)
    shard = OracleShard16(K, V, 2.6, 0.1, lptr, words[i0:i1], z0[i0:i1], seed=100 + rank, limit=limit)
    sampler = ShardedSampler(shard)
    n_sweeps = 5
    sampler.sweep(n_sweeps)
    if limit == 0:
        # refused, then a 32-bit exchange, alternately; the fifth sweep's exchange is refused too and sweep() ends by
        # flushing its deltas in 32 bits, so the replicas are consistent when it returns
        assert shard.n16 == 0 and shard.skipped == (n_sweeps + 1) // 2
    else:
        assert shard.n16 == n_sweeps and shard.skipped == 0
    zs = [None] * world
    dist.all_gather_object(zs, (i0, i1, shard.state.z.copy()))
    z = np.empty_like(z0)
    for a, b_, seg in zs:
        z[a:b_] = seg
    mi = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi, (z, words), 1)
    assert np.array_equal(shard.state.c.mi, mi), "replica != counts recomputed from the gathered zz"
    assert np.array_equal(shard.state.c.mm, mi.sum(1)) and shard.state.c.mm.sum() == len(words)
    dist.barrier()
    dist.destroy_process_group()


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bias_jl_b200 as b
    from bias_jl_b200.sharding import ShardedSampler, shard_csr

    K, V = 10, 25
    xx, _, _ = b.data.bars_lda_corpus(n_groups=41, n_tokens=30, seed=3)
    words = np.concatenate(xx)
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in xx])]).astype(np.int64)
    z0 = np.random.default_rng(0).integers(0, K, len(words))
    lo, hi, i0, i1, lptr = shard_csr(ptr, rank, world)
    shard = OracleShard(K, V, 2.5, 0.1, lptr, words[i0:i1], z0[i0:i1], seed=100 + rank)
    sampler = ShardedSampler(shard)
    # after construction every replica holds the GLOBAL counts
    mi0 = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi0, (z0, words), 1)
    assert np.array_equal(shard.state.c.mi, mi0), "initial merge of the local counts is wrong"
    sampler.sweep(4)
    zs = [None] * world
    dist.all_gather_object(zs, (i0, i1, shard.state.z.copy()))
    z = np.empty_like(z0)
    for a, b_, seg in zs:
        z[a:b_] = seg
    mi = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi, (z, words), 1)
    assert np.array_equal(shard.state.c.mi, mi), "replica != counts recomputed from the gathered zz"
    assert np.array_equal(shard.state.c.mm, mi.sum(1)) and shard.state.c.mm.sum() == len(words)
    # the local document-topic counts were never exchanged and still match the local zz
    nn = np.zeros_like(shard.state.nn)
    for j in range(len(lptr) - 1):
        nn[j] = np.bincount(shard.state.z[lptr[j]:lptr[j + 1]], minlength=K)
    assert np.array_equal(shard.state.nn, nn)
    moved = int((z != z0).sum())
    assert moved > 0
    np.save(os.path.join(out_dir, f"mi_{rank}.npy"), shard.state.c.mi)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(180)
def test_two_rank_exchange_over_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b_ = np.load(tmp_path / "mi_0.npy"), np.load(tmp_path / "mi_1.npy")
    assert np.array_equal(a, b_) and a.sum() == 41 * 30


def test_single_rank_needs_no_process_group():
    sys.path.insert(0, ROOT)
    import bias_jl_b200 as b
    from bias_jl_b200.sharding import ShardedSampler
    xx, _, _ = b.data.bars_lda_corpus(n_groups=5, n_tokens=20, seed=1)
    words = np.concatenate(xx)
    ptr = np.arange(6, dtype=np.int64) * 20
    z0 = np.random.default_rng(0).integers(0, 10, 100)
    shard = OracleShard(10, 25, 2.5, 0.1, ptr, words, z0, seed=1)
    ShardedSampler(shard).sweep(2)
    assert shard.state.c.mm.sum() == 100


@pytest.mark.timeout(180)
@pytest.mark.parametrize("limit", [32767, 0])
def test_two_rank_16_bit_exchange_over_gloo(tmp_path, limit):
    world = 2
    mp.spawn(_worker16, args=(world, _free_port(), str(tmp_path), limit), nprocs=world, join=True)
