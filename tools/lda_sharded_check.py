#!/usr/bin/env python
"""Sharded LDA with the in-library exchange, one process per rank:

    python -m torch.distributed.run --nproc-per-node N tools/lda_sharded_check.py

The ranks map each other's count tables with CUDA IPC (bias_model_share_*; csrc/comm.cu) and bias_model_sweep is
collective.  torch.distributed carries only host-side plumbing (handles, and the gather of the assignments for the
check).  Every rank needs its own GPU: the ranks' barrier kernels wait on one another and must not share a device.
After the sweeps every rank's replica of n_wk / n_k must equal the counts recomputed from the gathered
assignments (exact integers), the replicas' digests must agree, and the training log-likelihood must have improved.
One JSON line on rank 0.  LDA_ZIPF=1 makes the word frequencies heavy-tailed enough for hot words (int32 rows)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bias_jl_b200 as b
    from bias_jl_b200.sharding import attach_ranks, device_tensor, shard_bounds

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    n_gpu = torch.cuda.device_count()
    if world > n_gpu:
        raise SystemExit(f"{world} ranks on {n_gpu} GPU(s): every rank needs its own GPU (use bias_group for several shards on one)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # gloo: the collectives here are host plumbing and checks only
        dist.init_process_group("gloo")
    K, V, D, L = int(os.environ.get("LDA_K", 1024)), int(os.environ.get("LDA_V", 20000)), int(os.environ.get("LDA_DOCS", 40000)), 100
    n_sweeps = int(os.environ.get("LDA_SWEEPS", 6))
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    power = 6.0 if os.environ.get("LDA_ZIPF") else 2.0
    w_all = (torch.rand(D * L, device=dev, generator=g) ** power * V).to(torch.int32).clamp_(0, V - 1)      # same on every rank
    z_all = torch.randint(0, K, (D * L,), device=dev, dtype=torch.int32, generator=g)
    d0, d1 = shard_bounds(D, rank, world)
    words, z0 = w_all[d0 * L:d1 * L].cpu().numpy(), z_all[d0 * L:d1 * L].cpu().numpy()
    ptr = np.arange(d1 - d0 + 1, dtype=np.int64) * L
    ctx = b.Context(local, seed=500)
    lda = b.LDA(b.MultinomialDirichlet(V, 0.1 * V), K, 0.1)
    data = b.FlatData.from_csr(b._lib.INT, len(words), group_ptr=ptr, word=words)
    rs = b.ResidentSampler(ctx, lda, data, z0)
    if world > 1:
        attach_ranks(rs, len(words))
    rs.sweep(1)
    ll0 = rs.log_likelihood()[0]
    rs.sweep(n_sweeps - 1)
    ll1 = rs.log_likelihood()[0]
    cur, n_hot, _ = rs.narrow_info()
    z_loc = torch.from_numpy(rs.get_zz())
    digest = rs.replica_digest()
    if world > 1:
        zs = [None] * world
        dist.all_gather_object(zs, z_loc.numpy())
        z = torch.from_numpy(np.concatenate(zs)).to(dev)
        ds = [None] * world
        dist.all_gather_object(ds, digest)
    else:
        z, ds = z_loc.to(dev), [digest]
    rs.delta_begin_zero()                     # zero base => the delta buffer is the live [n_wk | n_k | n_items]
    (p32, n32, _), = rs.delta_pack()
    live = device_tensor(p32, n32, "int32", dev)
    Kpad = 128
    while Kpad < K:
        Kpad *= 2
    ref = torch.bincount(w_all.to(torch.int64) * Kpad + z.to(torch.int64), minlength=V * Kpad)
    ok = torch.equal(live[:V * Kpad].to(torch.int64), ref)
    ok = ok and torch.equal(live[V * Kpad:V * Kpad + Kpad].to(torch.int64), torch.bincount(z.to(torch.int64), minlength=Kpad))
    flags = [bool(ok), bool(ll1 > ll0), len(set(ds)) == 1, bool(cur)]
    if world > 1:
        allf = [None] * world
        dist.all_gather_object(allf, flags)
        flags = [all(f[i] for f in allf) for i in range(4)]
    if rank == 0:
        print(json.dumps({"world": world, "gpus": n_gpu, "tokens": D * L, "replicas_equal_recount": flags[0],
                          "loglik_improved": flags[1], "digests_equal": flags[2], "narrow_tables_current": flags[3],
                          "hot_words": n_hot, "max_word_freq": int(torch.bincount(w_all.to(torch.int64)).max()),
                          "moved_last": int(rs.stats().n_moved)}))
    rs.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
