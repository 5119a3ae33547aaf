"""GPU parity tests for the recurrent-CRP sampler (reference src/RCRP.jl:77-416): conditionals of the first,
middle and last epoch blocks on identical count states, whole sweeps, chain splits, hyper-parameters."""
import numpy as np
import pytest

import bias_jl_b200 as b
from oracle import oracle as orc
from helpers import assert_draws_match, norm_logp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0, seed=2468)
    yield c
    c.close()


def _rcrp_case(kind, seed, TT=5, Kmax=24):
    rng = np.random.default_rng(seed)
    lens = rng.integers(20, 45, TT)
    lens[1] = 0 if kind == "f64-empty" else lens[1]
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    N, KK = int(ptr[-1]), 5
    z0 = rng.integers(0, KK, N).astype(np.int32)
    t_mid = TT // 2
    seg = z0[ptr[t_mid]:ptr[t_mid + 1]]
    seg[seg == 3] = 0                                  # chain 3 has a gap at the middle epoch
    z0[:ptr[TT - 1]][z0[:ptr[TT - 1]] == 4] = 1        # chain 4 lives in the last epoch only
    aa = 0.3 + rng.random(TT)
    if kind.startswith("f64"):
        x = 10.0 * (z0 + 1) + 0.3 * rng.standard_normal(N)
        comp = b.Gaussian1DGaussian1D(float(x.mean()), 4.0, 0.05)
        odata, ocomp = orc.Data.reals(x), orc.Comps.g1g1(Kmax, float(x.mean()), 4.0, 0.05)
        xx = [x[ptr[t]:ptr[t + 1]] for t in range(TT)]
    elif kind == "int":
        V = 25
        w = rng.integers(0, V, N)
        comp = b.MultinomialDirichlet(V, 2.5)
        odata, ocomp = orc.Data.words(w), orc.Comps.md(Kmax, V, 2.5)
        xx = [w[ptr[t]:ptr[t + 1]] for t in range(TT)]
    else:
        V = 36
        sents, _, _ = b.data.bars_bmm_sentences(N, 12, 6, V, seed)
        comp = b.MultinomialDirichlet(V, 1.0)
        sp = np.concatenate([[0], np.cumsum([len(s) for s in sents])])
        odata = orc.Data.sents(sp, np.concatenate([s.w for s in sents]), np.concatenate([s.c for s in sents]))
        ocomp = orc.Comps.md(Kmax, V, 1.0)
        xx = [sents[ptr[t]:ptr[t + 1]] for t in range(TT)]
    rcrp = b.RCRP(comp, KK, aa, None, 1.0, 1.0, Kmax=Kmax)
    return rcrp, xx, odata, ocomp, ptr, z0


@pytest.mark.parametrize("kind", ["f64", "f64-empty", "int", "sent"])
def test_rcrp_conditionals_match_oracle(ctx, kind):
    rcrp, xx, odata, ocomp, ptr, z0 = _rcrp_case(kind, 11)
    data = b.FlatData(rcrp.component, xx, True)
    rs = b.ResidentSampler(ctx, rcrp, data, z0)
    st = orc.RCRPState(ocomp, odata, ptr, z0, rcrp.KK, rcrp.aa, rcrp.a1, rcrp.a2)
    N = data.n_items
    u = np.random.default_rng(5).random(N)
    raw, prob, draw = rs.conditionals(np.arange(N), u)
    assert raw.shape == (N, rcrp.KK + 1)
    grp = np.searchsorted(ptr, np.arange(N), side="right") - 1
    assert set(grp) >= {0, 2, len(ptr) - 2}            # first, middle and last epoch blocks are all exercised
    oraw, oprob, odraw = np.zeros_like(raw), np.zeros_like(prob), np.zeros(N, dtype=np.int64)
    for i in range(N):
        odraw[i], oraw[i], oprob[i] = st.conditional(int(grp[i]), i, float(u[i]))
    # chains that are not alive in the window have probability exactly 0 on both sides (pp = -Inf)
    dead = ~np.isfinite(oraw)
    assert dead.any() and np.array_equal(dead, ~np.isfinite(raw))
    assert (prob[dead] == 0).all()
    # parity level (a): normalised log-probabilities (the kernel drops the k-independent sum(L1) - L2)
    a, r = norm_logp(np.where(dead, -np.inf, raw)), norm_logp(oraw)
    err = np.abs(a[~dead] - r[~dead])
    assert (err <= 1e-5 * np.abs(r[~dead]) + 1e-6).all(), err.max()
    assert np.allclose(prob, oprob, rtol=1e-8, atol=1e-12)
    # parity level (b)
    assert_draws_match(draw, odraw, oprob, u, 1e-9, f"RCRP {kind} verification draws")
    rs.close()


@pytest.mark.parametrize("kind", ["f64", "f64-empty", "int", "sent"])
def test_rcrp_sweeps_keep_state_consistent(ctx, kind):
    rcrp, xx, odata, ocomp, ptr, z0 = _rcrp_case(kind, 21)
    data = b.FlatData(rcrp.component, xx, True)
    rs = b.ResidentSampler(ctx, rcrp, data, z0, sample_hyperparam=True, n_internals=4)
    aa_prev = np.array(rcrp.aa)
    for s in range(8):
        if s == 5:
            rs.set_sample_hyperparam(False)
        rs.sweep(1)
        rp, aa = rs.rcrp_state()
        z = rs.get_zz()
        KK = rs.stats().KK
        assert KK == rp.KK >= 1 and ((z >= 0) & (z < KK)).all()
        assert set(np.unique(z)) == set(range(KK))
        assert (aa > 0).all()
        assert (s < 5) == (not np.array_equal(aa, aa_prev))             # resampled only while switched on
        aa_prev = aa
        comp = rs.components()
        assert np.array_equal(comp["nn"], np.bincount(z, minlength=KK))
        nn = rs.group_counts()
        for t in range(len(ptr) - 1):
            assert np.array_equal(nn[t], np.bincount(z[ptr[t]:ptr[t + 1]], minlength=KK))
        # no chain has a gap: empty at t, alive before t and at t+1 (src/RCRP.jl:309-331) -- holds for the epochs
        # resampled before the split pass of the same sweep touched them again
        if kind.startswith("f64"):
            x = odata.x
            for k in range(KK):
                assert abs(comp["mu"][k] - x[z == k].mean()) < 1e-9
    if kind == "int":
        mi = np.zeros((KK, rcrp.component.dd), dtype=np.int64)
        np.add.at(mi, (z, np.asarray(odata.w)), 1)
        assert np.array_equal(comp["mi"], mi)
    rs.close()


def test_rcrp_chain_split(ctx):
    TT, n = 4, 30
    rng = np.random.default_rng(0)
    x = np.concatenate([np.full(n, 10.0), np.full(n, 20.0), np.full(n, 10.0), np.full(n, 10.0)])
    x = x + 0.01 * rng.standard_normal(len(x))
    xx = [x[t * n:(t + 1) * n] for t in range(TT)]
    zz = [np.zeros(n, dtype=np.int64), np.ones(n, dtype=np.int64), np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)]
    rcrp = b.RCRP(b.Gaussian1DGaussian1D(15.0, 100.0, 0.01), 2, 0.01, TT, 1.0, 1.0, Kmax=12)
    KK_list, KK_dict = b.RCRP_gibbs_sampler(rcrp, xx, zz, 0, 0, 1, False, 1, ctx=ctx)
    assert len(set(zz[0])) == 1 and len(set(zz[2])) == 1 and zz[0][0] != zz[2][0], "the chain was not split across the gap"
    assert rcrp.KK == KK_list[-1] == 3 and set(KK_dict) == {3}


def test_rcrp_gaussian_recovers_clusters_like_the_serial_sampler(ctx):
    """demo_RCRP_Gaussian1DGaussian1D.jl at reduced size: atoms N(10k, 0.01), KK starts at 1."""
    TT, n = 8, 200
    _, zz_true, _ = b.data.gen_rcrp_data([n] * TT, 0.5, seed=6)
    rng = np.random.default_rng(7)
    xx = [10.0 * (z + 1) + 0.1 * rng.standard_normal(len(z)) for z in zz_true]
    zt = np.concatenate(zz_true)
    q0 = b.Gaussian1DGaussian1D(float(xx[0].mean()), 2.0, 0.01)
    rcrp = b.RCRP(q0, 1, 1.0, TT, 1.0, 1.0, Kmax=64)
    zz = [np.zeros(n, dtype=np.int64) for _ in range(TT)]
    KK_list, KK_dict = b.RCRP_gibbs_sampler(rcrp, xx, zz, 10, 0, 10, True, 10, ctx=ctx)
    assert len(KK_list) == 10 and rcrp.KK == KK_list[-1] and rcrp.KK in KK_dict
    z = np.concatenate(zz)
    homog = sum(np.bincount(zt[z == c]).max() for c in np.unique(z)) / len(z)
    assert homog > 0.98, homog
    # serial oracle chain on the same data
    x = np.concatenate(xx)
    ptr = np.arange(TT + 1, dtype=np.int64) * n
    st = orc.RCRPState(orc.Comps.g1g1(64, float(xx[0].mean()), 2.0, 0.01), orc.Data.reals(x), ptr,
                       np.zeros(len(x), dtype=np.int64), 1, 1.0, 1.0, 1.0)
    r = orc.Rng(8)
    for s in range(20):
        assert not np.isnan(st.sweep(r, shuffle=True, sample_hyper=s < 10))
    n_true = len(np.unique(zt))
    assert n_true <= rcrp.KK <= st.KK + 4 and st.KK >= n_true, (rcrp.KK, st.KK, n_true)
    pos, nn = b.posterior(rcrp, xx, KK_dict, rcrp.KK, ctx=ctx)
    assert len(pos) == rcrp.KK and nn.shape == (TT, rcrp.KK) and nn.sum() == TT * n
    centres = sorted(round(p.mu / 10.0) for p in pos)
    assert set(centres) == set(int(k) + 1 for k in np.unique(zt))


def test_rcrp_long_birth_queue(ctx):
    """Epochs of 700 points starting from ONE component: the first sweep queues more new-component draws per epoch than
    the serial birth walk takes at once (256), so the parallel re-sampling rounds of the device loop run too."""
    TT, n = 6, 700
    rng = np.random.default_rng(3)
    zt = [rng.integers(t // 2, t // 2 + 4, n) for t in range(TT)]
    xx = [10.0 * (z + 1) + 0.1 * rng.standard_normal(n) for z in zt]
    allx = np.concatenate(xx)
    rcrp = b.RCRP(b.Gaussian1DGaussian1D(float(allx.mean()), 4.0, 0.01), 1, 1.0, TT, 1.0, 1.0, Kmax=64)
    data = b.FlatData(rcrp.component, xx, True)
    rs = b.ResidentSampler(ctx, rcrp, data, np.zeros(TT * n, dtype=np.int32), sample_hyperparam=True, n_internals=5)
    ptr = np.arange(TT + 1) * n
    for s in range(6):
        rs.sweep(1)
        z, KK = rs.get_zz(), rs.stats().KK
        assert ((z >= 0) & (z < KK)).all() and set(np.unique(z)) == set(range(KK))
        comp, nn = rs.components(), rs.group_counts()
        assert np.array_equal(comp["nn"], np.bincount(z, minlength=KK))
        for t in range(TT):
            assert np.array_equal(nn[t], np.bincount(z[ptr[t]:ptr[t + 1]], minlength=KK))
        for k in range(KK):
            assert abs(comp["mu"][k] - allx[z == k].mean()) < 1e-9
    truth = np.concatenate(zt)
    homog = sum(np.bincount(truth[z == c]).max() for c in np.unique(z)) / len(z)
    assert homog > 0.98 and len(np.unique(truth)) <= KK <= len(np.unique(truth)) + 6, (homog, KK)
    rs.close()


def test_rcrp_concentration_draws_match_the_serial_sampler_in_distribution(ctx):
    """The per-epoch concentration update of the device loop (src/RCRP.jl:60-74, drawn at the end of rcrp_loop_kernel) against
    the oracle's: 3000 epochs that all use exactly one component, so every aa[t] is an independent draw of the same update."""
    from scipy.stats import ks_2samp
    TT, n = 3000, 2
    x = np.zeros(TT * n)
    xx = [x[t * n:(t + 1) * n] for t in range(TT)]
    comp = b.Gaussian1DGaussian1D(0.0, 1e12, 0.01)        # the prior predictive is ~1e-7 of the component's: no births
    rcrp = b.RCRP(comp, 1, 1.0, TT, 1.0, 1.0, Kmax=8)
    rs = b.ResidentSampler(ctx, rcrp, b.FlatData(comp, xx, True), np.zeros(TT * n, dtype=np.int32), sample_hyperparam=True, n_internals=6)
    rs.sweep(1)
    rp, aa_gpu = rs.rcrp_state()
    assert rp.KK == 1
    rs.close()
    ptr = np.arange(TT + 1, dtype=np.int64) * n
    st = orc.RCRPState(orc.Comps.g1g1(8, 0.0, 1e12, 0.01), orc.Data.reals(x), ptr, np.zeros(TT * n, dtype=np.int64), 1, 1.0, 1.0, 1.0)
    st.sweep(orc.Rng(5), shuffle=True, sample_hyper=True, n_internals=6)
    assert st.KK == 1
    assert (aa_gpu > 0).all() and not np.allclose(aa_gpu, 1.0)
    assert ks_2samp(aa_gpu, st.aa).pvalue > 1e-3, (aa_gpu.mean(), st.aa.mean())
    assert abs(aa_gpu.mean() - st.aa.mean()) < 5 * st.aa.std() / np.sqrt(TT) * np.sqrt(2)


def test_rcrp_argument_errors(ctx):
    x = [np.array([1.0, 2.0])]
    with pytest.raises(ValueError):
        b.RCRP_gibbs_sampler(b.RCRP(b.Gaussian1DGaussian1D(0.0, 1.0, 1.0), 1, 1.0, 1), x, [np.zeros(2, dtype=np.int64)],
                             1, 0, 1, ctx=ctx)                             # a single epoch
    rng = np.random.default_rng(0)
    xs = [np.concatenate([10.0 * k + 0.01 * rng.standard_normal(20) for k in range(6)]) for _ in range(2)]
    rc = b.RCRP(b.Gaussian1DGaussian1D(25.0, 100.0, 0.001), 1, 1.0, 2, Kmax=3)
    with pytest.raises(b.BiasError) as e:
        b.RCRP_gibbs_sampler(rc, xs, [np.zeros(120, dtype=np.int64) for _ in range(2)], 3, 0, 1, ctx=ctx)
    assert e.value.code == 5
