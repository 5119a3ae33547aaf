#!/usr/bin/env python
"""bias_lda_run_multi — the ONE call the Julia shim makes for collapsed_gibbs_sampler!(lda, xx, zz, ...) (src/LDA.jl:67-72) — on the
GPUs of this box, from ONE process, no torch.distributed and no Python on the data path:

    python tools/lda_run_multi_check.py [--gpus N] [--sweeps S] [--docs D]

The BASELINE config-3 corpus (K = 1024, V = 50,000, 200-token documents; generated with bench.py's generator) goes in as host
arrays, the assignments come back in zz.  The call is timed as a whole (shard upload, initialisation pass, sweeps with the
library's peer-memory exchange, download) and the library reports where the call spent its time (bias_lda_run_multi_timing: set-up, sweeps, download).  Checks: every assignment in range, tokens moved in every sweep, and the word-topic counts of the returned zz equal
the counts a single-GPU model builds from the same zz (i.e. zz is a complete, consistent state).  One JSON line."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0, help="0 = all visible")
    ap.add_argument("--sweeps", type=int, default=10)
    ap.add_argument("--docs", type=int, default=500_000)
    args = ap.parse_args()
    import torch
    import bench
    import bias_jl_b200 as b
    from bias_jl_b200 import _lib
    from bias_jl_b200._lib import lib, check, ptr

    n_gpu = args.gpus or torch.cuda.device_count()
    devices = list(range(n_gpu))
    bench.D_TOTAL = args.docs
    dev = torch.device("cuda", 0)
    cdf = bench.topic_cdf_torch(dev)
    words = bench.gen_docs_torch(cdf, 0, args.docs, dev).cpu().numpy().astype(np.int32)
    del cdf
    torch.cuda.empty_cache()
    K, V, L = bench.K, bench.V, bench.L_DOC
    n = len(words)
    gptr = np.arange(args.docs + 1, dtype=np.int64) * L
    rng = np.random.default_rng(7)
    z0 = rng.integers(0, K, n).astype(np.int32)
    lda = b.LDA(b.MultinomialDirichlet(V, bench.BETA * V), K, bench.ALPHA)
    data = b.FlatData.from_csr(_lib.INT, n, group_ptr=gptr, word=words)
    spec = lda.component.spec(data.datum)
    dv = _lib.as_i32(devices)

    def run(n_sweeps, zz):
        stats = (_lib.SweepStats * max(1, n_sweeps))()
        t0 = time.perf_counter()
        check(lib().bias_lda_run_multi(ptr(dv), len(dv), C.c_uint64(11), C.byref(spec), C.byref(data.c), data.n_groups,
                                       ptr(data.group_ptr), ptr(zz), lda.KK, lda.aa, n_sweeps, stats))
        return time.perf_counter() - t0, list(stats)[:n_sweeps]

    def timing():
        ms = np.zeros(3)
        check(lib().bias_lda_run_multi_timing(ptr(ms)))
        return ms

    run(1, z0.copy())                                   # first call: contexts, peer mappings, module load
    zz = z0.copy()
    t_all, stats = run(args.sweeps, zz)
    setup_ms, sweeps_ms, finish_ms = timing()           # measured inside the library, devices idle at both ends of each span
    assert zz.min() >= 0 and zz.max() < K
    assert all(s.n_moved > 0 for s in stats)
    assert not np.array_equal(zz, z0)
    # the returned assignments as a single-GPU model sees them: its counts must be those of zz
    ctx = b.Context(0, seed=3)
    sub = slice(0, min(n, 20_000_000))
    sub_docs = sub.stop // L
    d1 = b.FlatData.from_csr(_lib.INT, sub.stop, group_ptr=gptr[:sub_docs + 1], word=words[sub])
    rs = b.ResidentSampler(ctx, lda, d1, zz[sub])
    mm = rs.components()["mm"]
    assert np.array_equal(mm, np.bincount(zz[sub], minlength=K)), "counts of the returned assignments"
    rs.close()
    ctx.close()
    per_sweep = 1e-3 * sweeps_ms / args.sweeps
    print(json.dumps({"entry_point": "bias_lda_run_multi", "process": "one process, no torch.distributed", "gpus": n_gpu, "tokens": int(n),
                      "K": K, "V": V, "sweeps": args.sweeps, "call_s": t_all, "setup_ms": setup_ms, "sweeps_ms": sweeps_ms, "finish_ms": finish_ms,
                      "ms_per_sweep": 1e3 * per_sweep, "tokens_per_s_sweeps": n / per_sweep,
                      "tokens_per_s_whole_call": n * args.sweeps / t_all, "moved_last": int(stats[-1].n_moved),
                      "checks": "assignments in range, tokens moved in every sweep, counts rebuilt from the returned zz consistent"}))


if __name__ == "__main__":
    main()
