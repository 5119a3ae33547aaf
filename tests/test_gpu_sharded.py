"""The sharded LDA sweep inside the library (csrc/comm.cu): documents split over ranks, replicas of the narrow count
tables reconciled once per sweep by the library's own reduce-scatter / all-gather kernel over peer memory.

* several shards in ONE process (bias_group; here all on one GPU, ordered by CUDA events);
* bias_lda_run_multi, the one-call entry point the Julia shim uses;
* one PROCESS per rank with CUDA IPC and the arrival-counter barriers (tools/lda_sharded_check.py under torchrun): needs
  one GPU per rank — kernels of different processes that wait on one another must not share a GPU (nothing guarantees
  that they run at the same time), so on a one-GPU box this test is skipped and the path is covered by bench.py at
  N >= 2, which ends with the same recount of every replica ("replica_check": "ok").
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import bias_jl_b200 as b
from bias_jl_b200.sharding import shard_csr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0, seed=31337)
    yield c
    c.close()


def _corpus(K, V, D, L, hot, seed):
    rng = np.random.default_rng(seed)
    words = rng.integers(0, V, D * L)
    if hot:
        # two words whose GLOBAL frequency passes 65,535 while no shard of four sees that many of either
        pos = rng.choice(D * L, 150_000, replace=False)
        words[pos[:80_000]] = 1
        words[pos[80_000:]] = V - 2
    ptr = np.arange(D + 1, dtype=np.int64) * L
    return ptr, words.astype(np.int64), rng.integers(0, K, D * L).astype(np.int32)


@pytest.mark.parametrize("K,V,D,L,world,hot", [(1024, 700, 2400, 100, 4, True), (64, 300, 900, 57, 3, False), (16, 40, 64, 33, 8, False)])
def test_group_of_shards_on_one_gpu(ctx, K, V, D, L, world, hot):
    ptr, words, z0 = _corpus(K, V, D, L, hot, seed=K + world)
    lda = b.LDA(b.MultinomialDirichlet(V, 0.1 * V), K, 0.1)
    shards, spans = [], []
    for r in range(world):
        lo, hi, i0, i1, lptr = shard_csr(ptr, r, world)
        data = b.FlatData.from_csr(b._lib.INT, i1 - i0, group_ptr=lptr, word=words[i0:i1].astype(np.int32))
        shards.append(b.ResidentSampler(ctx, lda, data, z0[i0:i1]))
        spans.append((i0, i1))
    grp = b.Group(shards)
    moved = 0
    for sweep in range(4):
        grp.sweep(1)
        moved += sum(s.stats().n_moved for s in shards)
        z = np.concatenate([s.get_zz() for s in shards])
        mi = np.zeros((K, V), dtype=np.int64)
        np.add.at(mi, (z, words), 1)
        digests = {s.replica_digest() for s in shards}
        assert len(digests) == 1, f"sweep {sweep}: the replicas differ"
        for s in shards:
            comp = s.components()
            assert np.array_equal(comp["mi"], mi), f"sweep {sweep}: a replica differs from the recount"
            assert np.array_equal(comp["mm"], mi.sum(1))
    assert moved > 0
    cur, n_hot, cap = shards[0].narrow_info()
    assert n_hot == (2 if hot else 0) and cap >= n_hot
    # a reload of every shard (new assignments) is merged by the next group sweep
    z1 = np.random.default_rng(1).integers(0, K, len(words)).astype(np.int32)
    for s, (i0, i1) in zip(shards, spans):
        s.reload(s.data, z1[i0:i1])
    grp.sweep(2)
    z = np.concatenate([s.get_zz() for s in shards])
    mi = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi, (z, words), 1)
    for s in shards:
        assert np.array_equal(s.components()["mi"], mi)
    with pytest.raises(b.BiasError):
        shards[0].sweep(1)                     # a shard of a group is swept through the group
    grp.close()
    for s in shards:
        s.close()


def test_lda_run_multi_matches_the_reference_entry_point_contract(ctx):
    """bias_lda_run_multi: zz mutated in place, counts of the returned assignments conserved, bars recovered as well as
    the single-GPU entry point does (two shards, both on device 0 here)."""
    K, V = 10, 25
    xx, true_zz, topics = b.data.bars_lda_corpus(n_groups=300, n_tokens=100, seed=2)
    lda = b.LDA(b.MultinomialDirichlet(V, 0.1 * V), K, 0.1)
    zz = [np.zeros(len(x), dtype=np.int64) for x in xx]
    b.init_zz(lda, zz, np.random.default_rng(0))
    before = [z.copy() for z in zz]
    stats = b.lda_run_multi([0, 0], lda, xx, zz, 120, seed=9)
    assert len(stats) == 120 and stats[-1].n_moved >= 0 and stats[0].n_moved > 0
    assert any((a != c).any() for a, c in zip(before, zz))
    z, w = np.concatenate(zz), np.concatenate(xx)
    assert ((z >= 0) & (z < K)).all()
    mi = np.zeros((K, V))
    np.add.at(mi, (z, w), 1)
    est = (mi + 0.1) / (mi + 0.1).sum(1, keepdims=True)
    got = sum(np.min(np.abs(est - t).sum(1)) < 0.25 for t in topics)
    assert got >= 6, got                        # a chain may keep one or two pairs of bars merged (tools/bars_modes.py)
    with pytest.raises(b.BiasError):
        b.lda_run_multi([0] * 9, lda, xx, zz, 1)


@pytest.mark.parametrize("zipf", [False, True])
def test_one_process_per_rank_over_cuda_ipc(zipf):
    """Two processes; each maps the other's tables with CUDA IPC, every exchange is ordered by the arrival counters in
    peer memory.  The check script recounts every replica from the gathered assignments."""
    n_gpu = b.device_count()
    if n_gpu < 2:
        pytest.skip("needs one GPU per rank (two)")
    env = dict(os.environ, LDA_DOCS="6000", LDA_K="256", LDA_V="3000", LDA_SWEEPS="4")
    if zipf:
        env["LDA_ZIPF"] = "1"
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(29541 + int(zipf)),
                          os.path.join(ROOT, "tools", "lda_sharded_check.py")],
                         capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["replicas_equal_recount"] and line["digests_equal"] and line["loglik_improved"] and line["narrow_tables_current"]
    assert line["world"] == 2 and line["gpus"] == n_gpu
    if zipf:
        assert line["hot_words"] >= 1
