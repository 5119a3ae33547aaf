#!/usr/bin/env python
"""Sharded BMM x MultinomialDirichlet x Sent (BASELINE config 2 shape) over NCCL, one process per GPU:

    python -m torch.distributed.run --nproc-per-node N tools/bmm_sharded_check.py

Sentences are split over the ranks; every rank keeps a replica of the component statistics and after each sweep the
count deltas are all-reduced (bias_model_delta_* + torch.distributed: sharding.ShardedSampler — the protocol of the models
that are not on the narrow-table LDA path).  After the sweeps every replica must equal the counts recomputed from the
gathered assignments.  One JSON line on rank 0."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bias_jl_b200 as b
    from bias_jl_b200.sharding import GpuShard, ShardedSampler, shard_items

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, V, L, n_sent = 64, 1024, 64, int(os.environ.get("BMM_SENTENCES", 1_000_000))
    sp, sw, sc, _, _ = b.data.bars_bmm_sentences_csr(n_sent, L, K, V, seed=123)          # same on every rank
    z_all = np.random.default_rng(0).integers(0, K, n_sent).astype(np.int32)
    i0, i1 = shard_items(n_sent, rank, world)
    lsp = (sp[i0:i1 + 1] - sp[i0]).astype(np.int64)
    lsw, lsc = sw[sp[i0]:sp[i1]], sc[sp[i0]:sp[i1]]
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = b.Context(local, seed=900 + rank, stream=stream.cuda_stream)
    bmm = b.BMM(b.MultinomialDirichlet(V, 1.0), K, 0.1)
    data = b.FlatData.from_csr(b._lib.SENT, i1 - i0, sent_ptr=lsp, sent_word=lsw, sent_count=lsc)
    rs = b.ResidentSampler(ctx, bmm, data, z_all[i0:i1])
    shard = GpuShard(rs, dev)
    if os.environ.get("BMM_FORCE32"):
        shard.delta_tensors16 = lambda world: None          # 32-bit deltas only
    sh = ShardedSampler(shard)
    sh.sweep(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n_timed = 5
    sh.sweep(n_timed)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = 1e3 * (time.perf_counter() - t0) / n_timed
    z_loc = rs.get_zz()
    if world > 1:
        zs = [None] * world
        dist.all_gather_object(zs, z_loc)
        z = np.concatenate(zs)
    else:
        z = z_loc
    comp = rs.components()
    item = np.repeat(np.arange(n_sent), np.diff(sp))
    mi = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi, (z[item], sw), sc)
    ok = bool(np.array_equal(comp["mi"], mi) and np.array_equal(comp["nn"], np.bincount(z, minlength=K)))
    if not ok:
        print(f"rank {rank}: mi differs in {int((comp['mi'] != mi).sum())} cells (max |diff| {int(np.abs(comp['mi'] - mi).max())}), "
              f"nn differs in {int((comp['nn'] != np.bincount(z, minlength=K)).sum())}", file=sys.stderr)
    t = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "sentences": n_sent, "tokens": n_sent * L, "ms_per_sweep": ms,
                          "sentences_per_s": n_sent / (ms * 1e-3), "replicas_equal_recount": bool(t.item()),
                          "moved_last": int(rs.stats().n_moved)}))
    rs.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
