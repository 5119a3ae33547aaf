#!/usr/bin/env python
"""Sharded LDA over NCCL (one process per GPU):  python -m torch.distributed.run --nproc-per-node N tools/lda_sharded_check.py
Documents are split over the ranks; after the sweeps every rank's replica of n_wk / n_k must equal the counts recomputed
from the gathered assignments (exact integers), and the training log-likelihood must have improved.  One JSON line on rank 0."""
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
    from bias_jl_b200.sharding import GpuShard, ShardedSampler, device_tensor, shard_bounds

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, V, D, L = 1024, 20000, int(os.environ.get("LDA_DOCS", 40000)), 100
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    w_all = (torch.rand(D * L, device=dev, generator=g) ** 2 * V).to(torch.int32).clamp_(0, V - 1)      # same on every rank
    z_all = torch.randint(0, K, (D * L,), device=dev, dtype=torch.int32, generator=g)
    d0, d1 = shard_bounds(D, rank, world)
    words, z0 = w_all[d0 * L:d1 * L].cpu().numpy(), z_all[d0 * L:d1 * L].cpu().numpy()
    ptr = np.arange(d1 - d0 + 1, dtype=np.int64) * L
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = b.Context(local, seed=500 + rank, stream=stream.cuda_stream)
    lda = b.LDA(b.MultinomialDirichlet(V, 0.1 * V), K, 0.1)
    data = b.FlatData.from_csr(b._lib.INT, len(words), group_ptr=ptr, word=words)
    rs = b.ResidentSampler(ctx, lda, data, z0)
    sh = ShardedSampler(GpuShard(rs, dev))
    ll0 = rs.log_likelihood()[0]
    sh.sweep(6)
    torch.cuda.synchronize()
    ll1 = rs.log_likelihood()[0]
    z_loc = torch.from_numpy(rs.get_zz()).to(dev)
    if world > 1:
        n_max = (D // world + 1) * L
        pad = torch.full((n_max,), -1, dtype=torch.int32, device=dev)
        pad[:len(z_loc)] = z_loc
        zs = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(zs, pad)
        z = torch.cat([t[t >= 0] for t in zs])
    else:
        z = z_loc
    rs.delta_begin_zero()                     # zero base => the delta buffer is the live [n_wk | n_k | n_items]
    (p32, n32, _), = rs.delta_pack()
    live = device_tensor(p32, n32, "int32", dev)
    Kpad = 1024
    ref = torch.bincount(w_all.to(torch.int64) * Kpad + z.to(torch.int64), minlength=V * Kpad)
    ok = torch.equal(live[:V * Kpad].to(torch.int64), ref)
    ok = ok and torch.equal(live[V * Kpad:V * Kpad + Kpad].to(torch.int64), torch.bincount(z.to(torch.int64), minlength=Kpad))
    t = torch.tensor([1 if ok else 0, 1 if ll1 > ll0 else 0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "tokens": D * L, "replicas_equal_recount": bool(t[0].item()),
                          "loglik_improved": bool(t[1].item()), "moved_last": int(rs.stats().n_moved)}))
    rs.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if int(t.min().item()) == 1 else 1)


if __name__ == "__main__":
    main()
