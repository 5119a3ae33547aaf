#!/usr/bin/env python
"""Prints the held-out perplexity trajectories of GPU chains (one GPU; 8 shards) next to the serial fixture chains of
tests/golden/lda_k1024_convergence.json (the numbers behind tests/test_gpu_convergence.py).  Needs a B200."""
import sys, json, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_convergence as t
import bias_jl_b200 as b
from bias_jl_b200.sharding import shard_csr
fx, (tptr, train, held, z0) = t._fixture()
sp = fx["spec"]; K, V = sp["K"], sp["V"]
ref = np.array([c["perplexity"] for c in fx["chains"]])
print("sweeps", fx["sweeps"]); print("serial mean", np.round(ref.mean(0), 1))
for world in (1, 8):
    for seed in (1, 2):
        ctx = b.Context(0, seed=1000 * seed + world)
        lda = b.LDA(b.MultinomialDirichlet(V, sp["beta"] * V), K, sp["alpha"])
        shards, hp = [], []
        for r in range(world):
            lo, hi, i0, i1, lptr = shard_csr(tptr, r, world)
            data = b.FlatData.from_csr(b._lib.INT, i1 - i0, group_ptr=lptr, word=train[i0:i1].astype(np.int32))
            shards.append(b.ResidentSampler(ctx, lda, data, z0[i0:i1].astype(np.int32)))
            hp.append(held[lo:hi])
        grp = b.Group(shards) if world > 1 else None
        got, done = [], 0
        for s in fx["sweeps"] + [120, 160]:
            (grp.sweep if grp else shards[0].sweep)(s - done); done = s
            got.append(t._perplexity(shards, hp))
        print(world, seed, np.round(got, 1))
        if grp: grp.close()
        for rs in shards: rs.close()
        ctx.close()
