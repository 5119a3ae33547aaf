#!/usr/bin/env python
"""Secondary measurements (not the bench.py line): sweep times of the other BASELINE configs
on one B200, with the serial oracle timed on a bounded sample next to each.

  C1  LDA bars K=10 V=25, 1k docs x 100 tokens
  C2  BMM x Sent, K=64, V=1024, 1M sentences x 64 tokens
  C4  HDP x Gaussian, groups of 100 points (size set by --hdp-groups)
  C5  RCRP x Gaussian, 100 epochs x 500 points (gen_RCRP_data, demo_RCRP_Gaussian1DGaussian1D.jl);
      RCRF x Gaussian, 100 epochs x 200 restaurants x 50 points
Writes one JSON object per config to stdout."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bias_jl_b200 as b          # noqa: E402
from oracle import oracle as orc  # noqa: E402  (the CPU sample next to each measurement: bench.py's cpu_baseline leg)


def timed_sweeps(rs, n, warm=2):
    rs.sweep(warm)
    rs.ctx.sync()
    t0 = time.perf_counter()
    rs.sweep(n)
    rs.ctx.sync()
    return (time.perf_counter() - t0) / n


def e2e(make_rs, n_items, n_sweeps=4):
    """End to end through the public API with host buffers: upload (model creation = the initialisation pass of the
    reference's sampler entry point), n_sweeps sweeps, assignments back on the host -> (items x sweeps / s, H2D, D2H bytes)."""
    rs = make_rs()          # warm: allocator, first-touch
    rs.sweep(1)
    rs.get_zz()
    rs.close()
    t0 = time.perf_counter()
    rs = make_rs()
    rs.sweep(n_sweeps)
    z = rs.get_zz()
    dt = time.perf_counter() - t0
    rs.close()
    return {"value": n_items * n_sweeps / dt, "unit": "items/s", "sweeps_per_call": n_sweeps, "d2h_bytes_per_step": int(z.nbytes)}


def c1(ctx):
    xx, _, _ = b.data.bars_lda_corpus(1000, 100, 10, 25, seed=123)
    lda = b.LDA(b.MultinomialDirichlet(25, 2.5), 10, 0.1)
    data = b.FlatData(lda.component, xx, True)
    z0 = np.random.default_rng(0).integers(0, 10, data.n_items).astype(np.int32)
    rs = b.ResidentSampler(ctx, lda, data, z0)
    t = timed_sweeps(rs, 50)
    rs.close()
    ee = e2e(lambda: b.ResidentSampler(ctx, lda, data, z0), data.n_items, 50)
    ee["h2d_bytes_per_step"] = int(data.n_items * 8 + data.group_ptr.nbytes)
    st = orc.LDAState(orc.Comps.md(10, 25, 2.5), orc.Data.words(np.concatenate(xx)), data.group_ptr, z0, 0.1, with_diag=False)
    u = np.random.default_rng(1).random(data.n_items)
    t0 = time.perf_counter()
    st.sweep(u, diag_mode=0)
    tc = time.perf_counter() - t0
    return {"config": "C1 LDA bars K=10 V=25 1000x100", "items": data.n_items, "gpu_ms_per_sweep": 1e3 * t,
            "gpu_items_per_s": data.n_items / t, "cpu_items_per_s": data.n_items / tc, "cpu_sample": "full corpus, 1 sweep", "e2e": ee,
            "alg_bytes_per_item": 4 * 128 + 12}


def c2(ctx, n_sent):
    K, V, L = 64, 1024, 64
    sp, sw, sc, _, _ = b.data.bars_bmm_sentences_csr(n_sent, L, K, V, seed=123)
    bmm = b.BMM(b.MultinomialDirichlet(V, 1.0), K, 0.1)
    data = b.FlatData.from_csr(b._lib.SENT, n_sent, sent_ptr=sp, sent_word=sw, sent_count=sc)
    z0 = np.random.default_rng(0).integers(0, K, n_sent).astype(np.int32)
    rs = b.ResidentSampler(ctx, bmm, data, z0)
    t = timed_sweeps(rs, 5, warm=2)
    moved = rs.stats().n_moved
    rs.close()
    ee = e2e(lambda: b.ResidentSampler(ctx, bmm, data, z0), n_sent, 4)
    ee["h2d_bytes_per_step"] = int(sp.nbytes + sw.nbytes + sc.nbytes + z0.nbytes)
    nnz = float(len(sw)) / n_sent
    ns = 300
    od = orc.Data.sents(sp[:ns + 1], sw[:sp[ns]], sc[:sp[ns]])
    st = orc.BMMState(orc.Comps.md(K, V, 1.0), od, z0[:ns], 0.1)
    u = np.random.default_rng(1).random(ns)
    t0 = time.perf_counter()
    st.sweep(u)
    tc = time.perf_counter() - t0
    return {"config": f"C2 BMM x Sent K=64 V=1024 {n_sent} sentences x 64 tokens", "items": n_sent, "tokens": n_sent * L,
            "gpu_ms_per_sweep": 1e3 * t, "gpu_items_per_s": n_sent / t, "gpu_tokens_per_s": n_sent * L / t,
            "n_moved_last": int(moved), "cpu_items_per_s": ns / tc,
            "cpu_sample": f"first {ns} sentences, 1 sweep of the serial sampler (dense O(V K) predictive as in the reference)", "e2e": ee,
            "alg_bytes_per_item": nnz * (8 + 4 * K) + 16}


def c4(ctx, n_groups):
    rng = np.random.default_rng(123)
    # CRF-like groups without the O(N^2) python seating loop: each group mixes a few of 20 atoms
    n_atoms = 20
    xx = []
    for _ in range(n_groups):
        ks = rng.permutation(n_atoms)[:int(rng.integers(1, 5))]
        z = ks[rng.integers(0, len(ks), 100)]
        xx.append(2.0 * (z + 1) + np.sqrt(0.001) * rng.standard_normal(100))
    allx = np.concatenate(xx)
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(allx.mean()), 10.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=256)
    ptr = np.arange(n_groups + 1, dtype=np.int64) * 100
    data = b.FlatData.from_csr(b._lib.F64, len(allx), group_ptr=ptr, x=allx)
    rs = b.ResidentSampler(ctx, hdp, data, np.zeros(len(allx), dtype=np.int32), sample_hyperparam=True, n_internals=10)
    ctx.sync()
    t0 = time.perf_counter()
    rs.sweep(3)
    ctx.sync()
    t_first = (time.perf_counter() - t0) / 3
    t = timed_sweeps(rs, 5, warm=2)
    s = rs.stats()
    rs.close()
    z00 = np.zeros(len(allx), dtype=np.int32)
    ee = e2e(lambda: b.ResidentSampler(ctx, hdp, data, z00, sample_hyperparam=True, n_internals=10), len(allx), 4)
    ee["h2d_bytes_per_step"] = int(allx.nbytes + z00.nbytes + ptr.nbytes)
    ng = min(n_groups, 40)
    st = orc.HDPState(orc.Comps.g1g1(256, float(allx.mean()), 10.0, 0.001), orc.Data.reals(allx[:ng * 100]), ptr[:ng + 1],
                      np.zeros(ng * 100, dtype=np.int64), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1)
    r = orc.Rng(1)
    tab = orc.stirling_table(128)
    for _ in range(3):
        st.sweep(r)
        st.resample_beta(tab, 128, 10, True, r)
    t0 = time.perf_counter()
    st.sweep(r)
    st.resample_beta(tab, 128, 10, True, r)
    tc = time.perf_counter() - t0
    return {"config": f"C4 HDP x Gaussian {n_groups} groups x 100 points, n_internals=10", "items": len(allx),
            "gpu_ms_per_sweep": 1e3 * t, "gpu_ms_per_sweep_first3": 1e3 * t_first, "gpu_items_per_s": len(allx) / t,
            "KK": int(s.KK), "cpu_items_per_s": ng * 100 / tc, "cpu_sample": f"first {ng} groups, 4th sweep of the serial sampler", "e2e": ee,
            "alg_bytes_per_item": 16}


def c4_crf(ctx, n_groups):
    """The same HDP data through the Chinese-restaurant-franchise sampler (CRF_gibbs_sampler!, src/HDP.jl:402-709)."""
    rng = np.random.default_rng(123)
    n_atoms = 20
    ks = rng.integers(0, n_atoms, (n_groups, 4))
    zt = np.take_along_axis(ks, rng.integers(0, 4, (n_groups, 100)), axis=1)
    allx = (2.0 * (zt + 1) + np.sqrt(0.001) * rng.standard_normal(zt.shape)).reshape(-1)
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(allx.mean()), 10.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=256)
    ptr = np.arange(n_groups + 1, dtype=np.int64) * 100
    data = b.FlatData.from_csr(b._lib.F64, len(allx), group_ptr=ptr, x=allx)
    rs = b.ResidentSampler(ctx, hdp, data, np.zeros(len(allx), dtype=np.int32), sample_hyperparam=True, n_internals=10, crf=True)
    ctx.sync()
    t0 = time.perf_counter()
    rs.sweep(3)
    ctx.sync()
    t_first = (time.perf_counter() - t0) / 3
    t = timed_sweeps(rs, 5, warm=2)
    KK = rs.stats().KK
    z = rs.get_zz()
    rs.close()
    homog = sum(np.bincount(zt.reshape(-1)[z == c]).max() for c in range(KK)) / len(z)
    ng = min(n_groups, 40)
    st = orc.CRFState(orc.Comps.g1g1(256, float(allx.mean()), 10.0, 0.001), orc.Data.reals(allx[:ng * 100]), ptr[:ng + 1],
                      np.zeros(ng * 100, dtype=np.int64), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1)
    r = orc.Rng(1)
    for _ in range(3):
        st.sweep(r, sample_hyper=True)
    t0 = time.perf_counter()
    st.sweep(r, sample_hyper=True)
    tc = time.perf_counter() - t0
    return {"config": f"C4 (CRF sampler) HDP x Gaussian {n_groups} groups x 100 points", "items": len(allx),
            "gpu_ms_per_sweep": 1e3 * t, "gpu_ms_per_sweep_first3": 1e3 * t_first, "gpu_items_per_s": len(allx) / t,
            "KK": int(KK), "homogeneity": homog, "cpu_items_per_s": ng * 100 / tc,
            "cpu_sample": f"first {ng} groups, 4th iteration of the serial sampler"}


def c5(ctx, n_epochs, n_per):
    _, zz_true, _ = b.data.gen_rcrp_data([n_per] * n_epochs, 0.5, seed=123)
    rng = np.random.default_rng(5)
    xx = [10.0 * (z + 1) + 0.1 * rng.standard_normal(len(z)) for z in zz_true]
    allx = np.concatenate(xx)
    rcrp = b.RCRP(b.Gaussian1DGaussian1D(float(xx[0].mean()), 2.0, 0.01), 1, 1.0, n_epochs, 1.0, 1.0, Kmax=256)
    data = b.FlatData(rcrp.component, xx, True)
    rs = b.ResidentSampler(ctx, rcrp, data, np.zeros(len(allx), dtype=np.int32), sample_hyperparam=True, n_internals=10)
    rs.sweep(10)                       # burn-in with hyper-parameter resampling (src/RCRP.jl:166-168)
    rs.set_sample_hyperparam(False)
    t = timed_sweeps(rs, 10, warm=2)
    KK = rs.stats().KK
    z = rs.get_zz()
    rs.close()
    z00 = np.zeros(len(allx), dtype=np.int32)
    ee = e2e(lambda: b.ResidentSampler(ctx, rcrp, data, z00, sample_hyperparam=True, n_internals=10), len(allx), 4)
    ee["h2d_bytes_per_step"] = int(allx.nbytes + z00.nbytes)
    zt = np.concatenate(zz_true)
    homog = sum(np.bincount(zt[z == c]).max() for c in np.unique(z)) / len(z)
    ptr = np.arange(n_epochs + 1, dtype=np.int64) * n_per
    st = orc.RCRPState(orc.Comps.g1g1(256, float(xx[0].mean()), 2.0, 0.01), orc.Data.reals(allx), ptr,
                       np.zeros(len(allx), dtype=np.int64), 1, 1.0, 1.0, 1.0)
    r = orc.Rng(1)
    for _ in range(3):
        st.sweep(r, shuffle=True, sample_hyper=True)
    t0 = time.perf_counter()
    st.sweep(r, shuffle=True)
    tc = time.perf_counter() - t0
    return {"config": f"C5 RCRP x Gaussian {n_epochs} epochs x {n_per} points", "items": len(allx),
            "gpu_ms_per_sweep": 1e3 * t, "gpu_items_per_s": len(allx) / t, "KK": int(KK), "KK_true": int(len(np.unique(zt))),
            "KK_serial": st.KK, "homogeneity": homog, "cpu_items_per_s": len(allx) / tc,
            "cpu_sample": "full data, 4th sweep of the serial sampler (O(n_{t+1,k}) look-ahead loops as in the reference)", "e2e": ee,
            "alg_bytes_per_item": 16}


def c5_rcrf(ctx, n_epochs, n_rest, n_per):
    """Config 5, second half: the recurrent Chinese restaurant franchise (RCRF_gibbs_sampler!, src/RCRF.jl) on
    streaming data: n_epochs epochs of n_rest restaurants, atoms drifting over the epochs."""
    rng = np.random.default_rng(321)
    G = n_epochs * n_rest
    eptr = np.arange(n_epochs + 1, dtype=np.int64) * n_rest
    ptr = np.arange(G + 1, dtype=np.int64) * n_per
    epoch = np.repeat(np.arange(G) // n_rest, n_per)
    atoms = (rng.integers(0, 4, G * n_per) + epoch // 5) % 24
    x = 3.0 * (atoms + 1) + 0.05 * rng.standard_normal(len(atoms))
    rcrf = b.RCRF(b.Gaussian1DGaussian1D(float(x.mean()), 100.0, 0.0025), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, n_epochs, Kmax=256)
    data = b.FlatData.from_csr(b._lib.F64, len(x), group_ptr=ptr, x=x)
    rs = b.ResidentSampler(ctx, rcrf, data, np.zeros(len(x), dtype=np.int32), sample_hyperparam=True, n_internals=10, epoch_ptr=eptr)
    rs.sweep(8)
    rs.set_sample_hyperparam(False)
    t = timed_sweeps(rs, 8, warm=2)
    KK = rs.stats().KK
    z = rs.get_zz()
    rs.close()
    homog = sum(np.bincount(atoms[z == c]).max() for c in np.unique(z)) / len(z)
    ne = min(n_epochs, 10)
    n_s = int(ptr[eptr[ne]])
    st = orc.RCRFState(orc.Comps.g1g1(256, float(x.mean()), 100.0, 0.0025), orc.Data.reals(x[:n_s]), eptr[:ne + 1], ptr[:eptr[ne] + 1],
                       np.zeros(n_s, dtype=np.int64), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1)
    r = orc.Rng(1)
    for _ in range(3):
        st.sweep(r, sample_hyper=True)
    t0 = time.perf_counter()
    st.sweep(r)
    tc = time.perf_counter() - t0
    return {"config": f"C5 RCRF x Gaussian {n_epochs} epochs x {n_rest} restaurants x {n_per} points", "items": len(x),
            "gpu_ms_per_sweep": 1e3 * t, "gpu_items_per_s": len(x) / t, "KK": int(KK), "atoms": int(len(np.unique(atoms))),
            "homogeneity": homog, "cpu_items_per_s": n_s / tc, "cpu_sample": f"first {ne} epochs, 4th iteration of the serial sampler"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sentences", type=int, default=1_000_000)
    ap.add_argument("--hdp-groups", type=int, default=100_000)
    ap.add_argument("--only", default="", help="comma-separated subset of c1,c2,c4,c4crf,c5,c5rcrf")
    args = ap.parse_args()
    ctx = b.Context(0, seed=123)
    runs = {"c1": lambda: c1(ctx), "c2": lambda: c2(ctx, args.sentences), "c4": lambda: c4(ctx, args.hdp_groups),
            "c4crf": lambda: c4_crf(ctx, args.hdp_groups), "c5": lambda: c5(ctx, 100, 500),
            "c5rcrf": lambda: c5_rcrf(ctx, 100, 200, 50)}
    for name, fn in runs.items():
        if args.only and name not in args.only.split(","):
            continue
        print(json.dumps(fn()), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
