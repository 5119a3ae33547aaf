"""CPU checks of the oracle's restatement of the recurrent-CRP sampler (reference src/RCRP.jl:77-416).
The reference has no tests and no golden vectors for this sampler ("parity unpinned");
these tests pin the oracle against an independent numpy restatement of the same formulas and against the
invariants the reference's loop maintains."""
import numpy as np
from scipy.special import gammaln

from oracle import oracle as orc


def _case(seed=0, TT=5, n=40, Kmax=24):
    rng = np.random.default_rng(seed)
    ptr = np.concatenate([[0], np.cumsum(rng.integers(n // 2, n, TT))]).astype(np.int64)
    N = int(ptr[-1])
    KK = 5
    x = 10.0 * rng.integers(1, KK + 1, N) + 0.1 * rng.standard_normal(N)
    z = rng.integers(0, KK, N)
    # make component 3 vanish from the middle epoch and component 4 exist only in the last one
    t_mid = TT // 2
    z[ptr[t_mid]:ptr[t_mid + 1]][z[ptr[t_mid]:ptr[t_mid + 1]] == 3] = 0
    z[:ptr[TT - 1]][z[:ptr[TT - 1]] == 4] = 1
    comps = orc.Comps.g1g1(Kmax, float(x.mean()), 4.0, 0.01)
    st = orc.RCRPState(comps, orc.Data.reals(x), ptr, z, KK, 0.7, 1.0, 1.0)
    return st, x, ptr


def _numpy_conditional(st, x, ptr, t, i):
    """log(cnt) + logpredictive + [L1_k(n_tk+1) - L1_k(n_tk)] with the L1 sums written as lgamma differences."""
    TT, KK = st.TT, st.KK
    nn = st.nn[:, :KK].copy()
    zi = int(st.z[i])
    nn[t, zi] -= 1
    # Gaussian statistics with the item removed
    z = st.z.copy()
    mask = np.ones(len(z), bool)
    mask[i] = False
    pp = np.full(KK + 1, -np.inf)
    m0, v0, vv = st.c.m0, st.c.v0, st.c.vv

    def logpred(k):
        sel = mask & (z == k) if k < KK else np.zeros(len(z), bool)
        n = sel.sum()
        if n:
            mu = x[sel].mean()
            pm = (n * v0 * mu + vv * m0) / (vv + n * v0)
            pv = v0 * vv / (vv + n * v0)
        else:
            pm, pv = m0, v0
        s2 = pv + vv
        return -0.5 * np.log(2 * np.pi * s2) - (x[i] - pm) ** 2 / (2 * s2)

    first, last = t == 0, t == TT - 1
    prev = np.zeros(KK, dtype=np.int64) if first else nn[t - 1]
    if last:
        for k in range(KK):
            if prev[k] + nn[t, k] > 0:
                pp[k] = np.log(prev[k] + nn[t, k]) + logpred(k)
        pp[KK] = np.log(st.aa[t]) + logpred(KK)
        return pp
    alive = (prev + nn[t] + nn[t + 1]) > 0
    a = st.aa[t] / (alive.sum() + 1)
    for k in range(KK):
        cnt = prev[k] + nn[t, k]
        if cnt > 0:
            nx, cur = nn[t + 1, k], nn[t, k]
            look = (gammaln(cur + 1 + a + nx) - gammaln(cur + 1 + a)) - (gammaln(cur + a + nx) - gammaln(cur + a))
            pp[k] = np.log(cnt) + logpred(k) + look
    pp[KK] = np.log(st.aa[t]) + logpred(KK)
    return pp


def _norm(raw):
    m = raw.max()
    return raw - (m + np.log(np.exp(raw - m).sum()))


def test_rcrp_conditionals_match_a_numpy_restatement():
    st, x, ptr = _case()
    rng = np.random.default_rng(1)
    grp = np.searchsorted(ptr, np.arange(ptr[-1]), side="right") - 1
    for i in rng.choice(int(ptr[-1]), 60, replace=False):
        t = int(grp[i])
        k, raw, prob = st.conditional(t, int(i), 0.5)
        ref = _numpy_conditional(st, x, ptr, t, int(i))
        assert np.array_equal(np.isfinite(raw), np.isfinite(ref)), (t, i)
        f = np.isfinite(ref)
        assert np.allclose(_norm(raw)[f], _norm(ref)[f], rtol=1e-9, atol=1e-9), (t, i)
        assert abs(prob.sum() - 1.0) < 1e-12 and (prob[~f] == 0).all()
        cdf = np.cumsum(prob)
        assert cdf[k] >= 0.5 and (k == 0 or cdf[k - 1] < 0.5)


def test_rcrp_conditional_leaves_state_untouched():
    st, x, ptr = _case(3)
    nn0, mu0, z0 = st.nn.copy(), st.c.mu.copy(), st.z.copy()
    st.conditional(2, int(ptr[2]) + 1, 0.3)
    assert np.array_equal(st.nn, nn0) and np.array_equal(st.c.mu, mu0) and np.array_equal(st.z, z0)


def _check_invariants(st, ptr):
    KK = st.KK
    z = st.z
    assert ((z >= 0) & (z < KK)).all()
    for t in range(st.TT):
        assert np.array_equal(st.nn[t, :KK], np.bincount(z[ptr[t]:ptr[t + 1]], minlength=KK))
    assert (st.nn[:, KK:] == 0).all()
    assert np.array_equal(st.c.nn[:KK], np.bincount(z, minlength=KK))
    assert (st.c.nn[:KK] > 0).all()                # empty chains are deleted at once (src/RCRP.jl:184-193)
    assert (st.c.nn[KK:] == 0).all()


def test_rcrp_sweeps_conserve_counts_and_split_chains():
    st, x, ptr = _case(5)
    r = orc.Rng(9)
    for s in range(6):
        ll = st.sweep(r, shuffle=bool(s % 2), sample_hyper=(s < 3), n_internals=4)
        assert not np.isnan(ll)
        _check_invariants(st, ptr)
        assert (st.aa > 0).all()
    # Gaussian running means agree with a recomputation from the labels
    for k in range(st.KK):
        assert abs(st.c.mu[k] - x[st.z == k].mean()) < 1e-9


def test_rcrp_chain_split_rule():
    """src/RCRP.jl:309-331: after a middle epoch t is resampled, a chain that is empty at t but alive before and
    at t+1 hands everything from t+1 on to a fresh component."""
    TT, n = 4, 30
    ptr = np.arange(TT + 1, dtype=np.int64) * n
    x = np.concatenate([np.full(n, 10.0), np.full(n, 20.0), np.full(n, 10.0), np.full(n, 10.0)])
    x = x + 0.01 * np.random.default_rng(0).standard_normal(len(x))
    z = np.concatenate([np.zeros(n), np.ones(n), np.zeros(n), np.zeros(n)]).astype(np.int64)
    comps = orc.Comps.g1g1(12, 15.0, 100.0, 0.01)
    st = orc.RCRPState(comps, orc.Data.reals(x), ptr, z, 2, 0.01, 1.0, 1.0)
    st.sweep(orc.Rng(1))
    # epoch 1 keeps its own atom (20), so chain 0 is empty there and epochs 2.. are re-labelled
    z = st.z
    assert len(set(z[:n])) == 1 and len(set(z[2 * n:])) == 1
    assert z[0] != z[2 * n], "the chain was not split across the gap"
    _check_invariants(st, ptr)


def test_rcrp_recovers_the_clusters_of_generated_data():
    import bias_jl_b200 as b
    nn_true, zz_true, KK_true = b.data.gen_rcrp_data([60] * 6, 0.5, seed=4)
    assert nn_true.shape == (6, KK_true) and nn_true.sum() == 360
    rng = np.random.default_rng(2)
    zt = np.concatenate(zz_true)
    x = 10.0 * (zt + 1) + 0.1 * rng.standard_normal(len(zt))
    ptr = np.arange(7, dtype=np.int64) * 60
    comps = orc.Comps.g1g1(64, float(x[:60].mean()), 2.0, 0.01)
    st = orc.RCRPState(comps, orc.Data.reals(x), ptr, np.zeros(len(x), dtype=np.int64), 1, 1.0, 1.0, 1.0)
    r = orc.Rng(3)
    for s in range(12):
        # the diagnostic is log(pdf(...)) (src/distributions.jl:37) and may underflow to -Inf early on; NaN = capacity
        assert not np.isnan(st.sweep(r, shuffle=True, sample_hyper=s < 5))
    homog = sum(np.bincount(zt[st.z == c]).max() for c in np.unique(st.z)) / len(zt)
    assert homog > 0.98
    assert len(np.unique(zt)) <= st.KK <= len(np.unique(zt)) + 8       # chain splits may duplicate an atom across gaps
