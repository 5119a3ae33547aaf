#!/usr/bin/env python
"""Sharded HDP over NCCL (one process per GPU):  python -m torch.distributed.run --nproc-per-node N tools/hdp_sharded_check.py
Groups are split over the ranks; after every sweep the replicas of KK / beta / alpha0 / gamma / component statistics must
be identical and equal the recomputation from the gathered assignments.  Prints one JSON line on rank 0."""
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
    from bias_jl_b200.sharding import DistComm, ShardedHDP, shard_csr

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_groups, n_per, n_atoms = int(os.environ.get("HDP_GROUPS", 20000)), 100, 20
    rng = np.random.default_rng(123)
    ks = rng.integers(0, n_atoms, (n_groups, 4))
    zt = np.take_along_axis(ks, rng.integers(0, 4, (n_groups, n_per)), axis=1)
    x = (2.0 * (zt + 1) + np.sqrt(0.001) * rng.standard_normal(zt.shape)).reshape(-1)
    ptr = np.arange(n_groups + 1, dtype=np.int64) * n_per
    lo, hi, i0, i1, lptr = shard_csr(ptr, rank, world)
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(x.mean()), 10.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=256)
    ctx = b.Context(local, seed=1000 + rank, stream=torch.cuda.current_stream().cuda_stream)
    data = b.FlatData.from_csr(b._lib.F64, i1 - i0, group_ptr=lptr, x=x[i0:i1])
    rs = b.ResidentSampler(ctx, hdp, data, np.zeros(i1 - i0, dtype=np.int32), sample_hyperparam=True, n_internals=10)
    sh = ShardedHDP([rs], DistComm(), dev, host_seed=99)
    n_sweeps = 12
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sh.sweep(n_sweeps)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n_sweeps
    hp, beta = rs.hdp_state()
    z_local = torch.from_numpy(rs.get_zz().astype(np.int64)).to(dev)
    sig = torch.tensor([hp.KK, hp.gg, hp.aa, float(beta.sum()), float((beta * np.arange(len(beta))).sum())],
                       dtype=torch.float64, device=dev)
    ok = True
    if world > 1:
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        ok = all(torch.equal(s, sigs[0]) for s in sigs)
        n_max = (n_groups // world + 1) * n_per
        pad = torch.full((n_max,), -1, dtype=torch.int64, device=dev)
        pad[:len(z_local)] = z_local
        zs = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(zs, pad)
        z = torch.cat([t[t >= 0] for t in zs]).cpu().numpy()
    else:
        z = z_local.cpu().numpy()
    comp = rs.components()
    ok = ok and np.array_equal(comp["nn"], np.bincount(z, minlength=hp.KK))
    sums = np.bincount(z, weights=x, minlength=hp.KK)
    ok = ok and np.allclose(comp["mu"] * comp["nn"], sums, rtol=1e-9)
    homog = sum(np.bincount(zt.reshape(-1)[z == c]).max() for c in range(hp.KK)) / len(z)
    ok_all = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "points": int(len(x)), "ms_per_sweep": 1e3 * dt, "points_per_s": len(x) / dt,
                          "KK": int(hp.KK), "atoms": n_atoms, "homogeneity": homog, "replicas_consistent": bool(ok_all.item())}))
    rs.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok_all.item() else 1)


if __name__ == "__main__":
    main()
