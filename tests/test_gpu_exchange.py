"""The multi-GPU exchange kernels on ONE GPU: two shards of a corpus live in two resident samplers on
the same device; their count deltas are summed (what the NCCL all-reduce does across ranks) and
applied.  Replicas must stay identical and equal the counts recomputed from the gathered zz."""
import numpy as np
import pytest
import torch

import bias_jl_b200 as b
from bias_jl_b200.sharding import device_tensor, shard_csr, shard_items

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0, seed=77)
    yield c
    c.close()


def _exchange(samplers, first=False):
    dev = torch.device("cuda", 0)
    if first:
        for rs in samplers:
            rs.delta_begin_zero()
    packs = [rs.delta_pack() for rs in samplers]
    samplers[0].ctx.sync()
    for seg in range(len(packs[0])):
        ts = [device_tensor(*p[seg], dev) for p in packs]
        total = ts[0].clone()
        for t in ts[1:]:
            total += t
        for t in ts:
            t.copy_(total)
    torch.cuda.synchronize()
    for rs in samplers:
        rs.delta_apply()


def test_two_lda_shards_on_one_gpu(ctx):
    K, V = 10, 25
    xx, _, _ = b.data.bars_lda_corpus(n_groups=61, n_tokens=40, seed=9)
    words = np.concatenate(xx)
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in xx])]).astype(np.int64)
    z0 = np.random.default_rng(1).integers(0, K, len(words)).astype(np.int32)
    lda = b.LDA(b.MultinomialDirichlet(V, 2.5), K, 0.1)
    samplers, spans = [], []
    for r in range(2):
        lo, hi, i0, i1, lptr = shard_csr(ptr, r, 2)
        data = b.FlatData.from_csr(b._lib.INT, i1 - i0, group_ptr=lptr, word=words[i0:i1])
        samplers.append(b.ResidentSampler(ctx, lda, data, z0[i0:i1]))
        spans.append((i0, i1))
    _exchange(samplers, first=True)
    mi0 = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi0, (z0, words), 1)
    for rs in samplers:
        assert np.array_equal(rs.components()["mi"], mi0)
    for _ in range(5):
        for rs in samplers:
            rs.sweep(1)
        _exchange(samplers)
    z = np.concatenate([rs.get_zz() for rs in samplers])
    mi = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi, (z, words), 1)
    assert (z != z0).any()
    for rs in samplers:
        comp = rs.components()
        assert np.array_equal(comp["mi"], mi) and np.array_equal(comp["mm"], mi.sum(1))
        rs.close()


def test_two_bmm_gaussian_shards_on_one_gpu(ctx):
    K = 5
    x, _, _ = b.data.gaussian_mixture(400, K, 0.01, seed=2)
    z0 = np.random.default_rng(3).integers(0, K, len(x)).astype(np.int32)
    bmm = b.BMM(b.Gaussian1DGaussian1D(float(x.mean()), 2.0, 0.01), K, 1.0)
    samplers = []
    for r in range(2):
        i0, i1 = shard_items(len(x), r, 2)
        data = b.FlatData.from_csr(b._lib.F64, i1 - i0, x=x[i0:i1])
        samplers.append(b.ResidentSampler(ctx, bmm, data, z0[i0:i1]))
    _exchange(samplers, first=True)
    for rs in samplers:
        assert np.array_equal(rs.components()["nn"], np.bincount(z0, minlength=K))
    for _ in range(5):
        for rs in samplers:
            rs.sweep(1)
        _exchange(samplers)
    z = np.concatenate([rs.get_zz() for rs in samplers])
    for rs in samplers:
        comp = rs.components()
        assert np.array_equal(comp["nn"], np.bincount(z, minlength=K))
        for k in range(K):
            if comp["nn"][k]:
                assert abs(comp["mu"][k] - x[z == k].mean()) < 1e-9
        rs.close()


def test_16_bit_exchange_of_word_topic_deltas(ctx):
    """bias_model_delta_pack16 / apply16: int16 deltas for n_wk, int32 for n_k / n_items, overflow flag raised when a
    delta times the number of ranks would not fit; both shards end up with the recount from the gathered zz."""
    K, V = 10, 25
    xx, _, _ = b.data.bars_lda_corpus(n_groups=50, n_tokens=40, seed=4)
    words = np.concatenate(xx)
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in xx])]).astype(np.int64)
    z0 = np.random.default_rng(2).integers(0, K, len(words)).astype(np.int32)
    lda = b.LDA(b.MultinomialDirichlet(V, 2.5), K, 0.1)
    dev = torch.device("cuda", 0)
    samplers = []
    for r in range(2):
        lo, hi, i0, i1, lptr = shard_csr(ptr, r, 2)
        data = b.FlatData.from_csr(b._lib.INT, i1 - i0, group_ptr=lptr, word=words[i0:i1])
        samplers.append(b.ResidentSampler(ctx, lda, data, z0[i0:i1]))

    def exchange16(world):
        packs = [rs.delta_pack16(world) for rs in samplers]
        ctx.sync()
        tens = [[device_tensor(*seg, dev) for seg in p] for p in packs]
        for seg in range(2):
            total = tens[0][seg].clone()
            total += tens[1][seg]
            for t in tens:
                t[seg].copy_(total)
        torch.cuda.synchronize()
        return int(tens[0][1][-4].item())

    for rs in samplers:
        rs.delta_begin_zero()
    assert exchange16(2) == 0                    # the initial merge: counts of a 2000-token corpus fit easily
    for rs in samplers:
        rs.delta_apply16()
    mi0 = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi0, (z0, words), 1)
    for rs in samplers:
        assert np.array_equal(rs.components()["mi"], mi0)
    for _ in range(3):
        for rs in samplers:
            rs.sweep(1)
        assert exchange16(2) == 0
        for rs in samplers:
            rs.delta_apply16()
    z = np.concatenate([rs.get_zz() for rs in samplers])
    mi = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi, (z, words), 1)
    for rs in samplers:
        comp = rs.components()
        assert np.array_equal(comp["mi"], mi) and np.array_equal(comp["mm"], mi.sum(1))
    # a "world" so large that any non-zero delta could overflow raises the flag; nothing is applied, and the 32-bit
    # exchange that follows gives the right tables
    for rs in samplers:
        rs.sweep(1)
    assert exchange16(40000) != 0
    for rs in samplers:
        rs.delta_apply16()                       # gated on the device: the raised flag makes this a no-op
    _exchange(samplers)
    z = np.concatenate([rs.get_zz() for rs in samplers])
    mi = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi, (z, words), 1)
    for rs in samplers:
        assert np.array_equal(rs.components()["mi"], mi)
        rs.close()


def test_sharded_bmm_over_nccl_two_processes():
    """BMM x MultinomialDirichlet x Sent at the config-2 shape on two GPUs through ShardedSampler (count deltas + NCCL
    all-reduce, 16-bit packed with the 32-bit fall-back): 1 M sentences move some cells by more than 16,383 in one
    sweep, so the overflow flag comes up, the refused exchange is completed in 32 bits (ShardedSampler.flush) and every
    replica equals the recount from the gathered assignments (tools/bmm_sharded_check.py)."""
    import json
    import os
    import subprocess
    import sys
    if b.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29537", os.path.join(root, "tools", "bmm_sharded_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["replicas_equal_recount"] and line["world"] == 2
