#!/usr/bin/env python
"""Number of components the direct-assignment HDP sampler ends with on the config-4 data (20 atoms) for several seeds:
a check that a kernel change has not changed what the chain finds.  python tools/hdp_kk_check.py [groups] [sweeps]"""
import json, sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bias_jl_b200 as b

n_groups = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = []
for seed in (1, 2, 3, 4, 5, 6):
    rng = np.random.default_rng(123)
    ks = rng.integers(0, 20, (n_groups, 4))
    zt = np.take_along_axis(ks, rng.integers(0, 4, (n_groups, 100)), axis=1)
    allx = (2.0 * (zt + 1) + np.sqrt(0.001) * rng.standard_normal(zt.shape)).reshape(-1)
    ctx = b.Context(0, seed=seed)
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(allx.mean()), 10.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=256)
    ptr = np.arange(n_groups + 1, dtype=np.int64) * 100
    data = b.FlatData.from_csr(b._lib.F64, len(allx), group_ptr=ptr, x=allx)
    rs = b.ResidentSampler(ctx, hdp, data, np.zeros(len(allx), dtype=np.int32), sample_hyperparam=True, n_internals=10)
    kk = []
    for s in range(sweeps):
        rs.sweep(1)
        kk.append(int(rs.stats().KK))
    z = rs.get_zz()
    homog = sum(np.bincount(zt.reshape(-1)[z == c]).max() for c in np.unique(z)) / len(z)
    out.append({"seed": seed, "KK": kk, "homogeneity": homog})
    rs.close()
    ctx.close()
print(json.dumps(out))
