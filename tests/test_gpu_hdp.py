"""GPU parity tests for the HDP direct-assignment sampler (reference src/HDP.jl:166-399):
Stirling table, table-count conditionals, K+1 assignment conditionals, and whole sweeps."""
import numpy as np
import pytest

import bias_jl_b200 as b
from bias_jl_b200.models import hdp_table_conditionals, stirling_rows
from oracle import oracle as orc
from helpers import assert_conditionals_close, assert_draws_match

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0, seed=4321)
    yield c
    c.close()


def test_stirling_table_matches_oracle(ctx):
    n = 300
    dev = stirling_rows(ctx, n)
    ref = orc.stirling_table(n)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(dev), fin)
    assert np.allclose(dev[fin], ref[fin], rtol=1e-12, atol=1e-9)
    row5 = np.exp(dev[10:15])
    assert np.allclose(row5 / row5.max(), np.array([24, 50, 35, 10, 1]) / 50.0, rtol=1e-12)
    # every row is max-normalised
    for r in (1, 2, 17, 300):
        assert dev[r * (r - 1) // 2: r * (r + 1) // 2].max() == 0.0


def test_table_count_conditionals_match_oracle(ctx):
    rng = np.random.default_rng(0)
    nmax = 400
    ref_tab = orc.stirling_table(nmax)
    n = np.concatenate([[1, 2, 3, nmax], rng.integers(1, nmax, 60)]).astype(np.int32)
    a = np.concatenate([[0.5, 3.0, 1e-3, 2.5], rng.gamma(1.0, 2.0, 60) + 1e-3])
    u = rng.random(len(n))
    prob, draws = hdp_table_conditionals(ctx, n, a, u)
    for q in range(len(n)):
        m_ref, p_ref = orc.hdp_table_conditional(ref_tab, int(n[q]), float(a[q]), float(u[q]))
        assert np.allclose(prob[q, :n[q]], p_ref, rtol=1e-9, atol=1e-300)
        assert (prob[q, n[q]:] == 0).all()
        if draws[q] != m_ref:
            assert np.min(np.abs(np.cumsum(p_ref) - u[q])) < 1e-9
    with pytest.raises(b.BiasError) as e:
        hdp_table_conditionals(ctx, [10001], [1.0], [0.5])          # the reference's table has 10k rows: BoundsError
    assert e.value.code == 5


def _hdp_case(kind, seed):
    rng = np.random.default_rng(seed)
    KK, Kmax = 4, 16
    ptr = np.array([0, 30, 30, 75, 120], dtype=np.int64)
    N = int(ptr[-1])
    if kind == "f64":
        x = np.repeat([2.0, 4.0, 6.0, 8.0], 30) + 0.05 * rng.standard_normal(N)
        rng.shuffle(x)
        comp = b.Gaussian1DGaussian1D(float(x.mean()), 10.0, 0.01)
        odata = orc.Data.reals(x)
        ocomp = orc.Comps.g1g1(Kmax, float(x.mean()), 10.0, 0.01)
        xx = [x[ptr[j]:ptr[j + 1]] for j in range(4)]
    elif kind == "int":
        V = 25
        w = rng.integers(0, V, N)
        comp = b.MultinomialDirichlet(V, 2.5)
        odata = orc.Data.words(w)
        ocomp = orc.Comps.md(Kmax, V, 2.5)
        xx = [w[ptr[j]:ptr[j + 1]] for j in range(4)]
    else:
        V = 25
        sents, true_zz, _ = b.data.bars_bmm_sentences(N, 12, 4, V, seed)
        comp = b.MultinomialDirichlet(V, 2.5)
        sp = np.concatenate([[0], np.cumsum([len(s) for s in sents])])
        odata = orc.Data.sents(sp, np.concatenate([s.w for s in sents]), np.concatenate([s.c for s in sents]))
        ocomp = orc.Comps.md(Kmax, V, 2.5)
        xx = [sents[ptr[j]:ptr[j + 1]] for j in range(4)]
    z0 = rng.integers(0, KK, N).astype(np.int32)
    hdp = b.HDP(comp, KK, 1.0, 0.1, 0.1, 1.5, 0.1, 0.1, Kmax=Kmax)
    return hdp, xx, odata, ocomp, ptr, z0


@pytest.mark.parametrize("kind", ["f64", "int", "sent"])
def test_assignment_conditionals_match_oracle(ctx, kind):
    hdp, xx, odata, ocomp, ptr, z0 = _hdp_case(kind, 3)
    data = b.FlatData(hdp.component, xx, True)
    rs = b.ResidentSampler(ctx, hdp, data, z0)
    st = orc.HDPState(ocomp, odata, ptr, z0, hdp.KK, hdp.gg, hdp.g1, hdp.g2, hdp.aa, hdp.a1, hdp.a2)
    N = data.n_items
    rng = np.random.default_rng(8)
    u = rng.random(N)
    raw, prob, draw = rs.conditionals(np.arange(N), u)
    assert raw.shape == (N, hdp.KK + 1)
    grp = np.searchsorted(ptr, np.arange(N), side="right") - 1
    oraw, oprob, odraw = np.zeros_like(raw), np.zeros_like(prob), np.zeros(N, dtype=np.int64)
    for i in range(N):
        odraw[i], oraw[i], oprob[i] = st.conditional(int(grp[i]), i, float(u[i]))
    assert_conditionals_close(raw, oraw, f"HDP {kind}")
    assert np.allclose(raw, oraw, rtol=1e-9, atol=1e-8)
    assert_draws_match(draw, odraw, oprob, u, 1e-9, "HDP verification draws")
    rs.close()


@pytest.mark.parametrize("kind", ["f64", "int", "sent"])
def test_sweeps_keep_state_consistent(ctx, kind):
    hdp, xx, odata, ocomp, ptr, z0 = _hdp_case(kind, 5)
    data = b.FlatData(hdp.component, xx, True)
    rs = b.ResidentSampler(ctx, hdp, data, z0, sample_hyperparam=True, n_internals=5)
    for _ in range(8):
        rs.sweep(1)
        s = rs.stats()
        hp, beta = rs.hdp_state()
        z = rs.get_zz()
        assert s.KK == hp.KK >= 1 and ((z >= 0) & (z < hp.KK)).all()
        assert set(np.unique(z)) == set(range(hp.KK))                 # every surviving component is non-empty
        assert len(beta) == hp.KK + 1 and abs(beta.sum() - 1.0) < 1e-9 and (beta > 0).all()
        assert hp.gg > 0 and hp.aa > 0
        comp = rs.components()
        assert np.array_equal(comp["nn"], np.bincount(z, minlength=hp.KK))
        nn = rs.group_counts()
        for j in range(len(ptr) - 1):
            assert np.array_equal(nn[j], np.bincount(z[ptr[j]:ptr[j + 1]], minlength=hp.KK))
    if kind != "f64":
        V = hdp.component.dd
        mi = np.zeros((hp.KK, V), dtype=np.int64)
        if kind == "int":
            np.add.at(mi, (z, np.asarray(odata.w)), 1)
        else:
            item = np.repeat(np.arange(data.n_items), np.diff(odata.sptr))
            np.add.at(mi, (z[item], odata.sw), odata.sc)
        assert np.array_equal(comp["mi"], mi)
    rs.close()


def test_hdp_gaussian_finds_the_number_of_atoms(ctx):
    """demo_HDP_Gaussian1DGaussian1D.jl: CRF-generated groups, atoms N(2k, 0.001); start from KK=1."""
    zz_true, n_atoms = b.data.gen_crf_data([100] * 30, 1.0, 0.5, seed=3)
    rng = np.random.default_rng(4)
    xx = [2.0 * (z + 1) + np.sqrt(0.001) * rng.standard_normal(len(z)) for z in zz_true]
    allx = np.concatenate(xx)
    make = lambda: b.HDP(b.Gaussian1DGaussian1D(float(allx.mean()), 10.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=64)
    hdp = make()
    zz = [np.zeros(len(x), dtype=np.int64) for x in xx]
    KK_list, KK_dict, betas, gammas, alphas = b.collapsed_gibbs_sampler(hdp, xx, zz, 20, 1, 10, True, 5, ctx=ctx)
    assert len(KK_list) == 10 and len(betas) == 10 and hdp.KK == KK_list[-1]
    used = np.unique(np.concatenate(zz_true))
    assert abs(hdp.KK - len(used)) <= 2, (hdp.KK, len(used))
    # the serial reference sampler on the same data lands on a similar number of components
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in xx])])
    Kmax = 64
    st = orc.HDPState(orc.Comps.g1g1(Kmax, float(allx.mean()), 10.0, 0.001), orc.Data.reals(allx), ptr,
                      np.zeros(len(allx), dtype=np.int64), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1)
    r = orc.Rng(5)
    tab = orc.stirling_table(128)
    for _ in range(15):
        assert not np.isnan(st.sweep(r))
        st.resample_beta(tab, 128, 5, True, r)
    # (the serial chain tends to carry a few more small components this early; the parallel one must not over-create)
    assert st.KK >= len(used) - 2 and hdp.KK <= st.KK + 2, (hdp.KK, st.KK, len(used))
    # every inferred component is (almost) pure in the generating labels; duplicates of one atom are
    # allowed this early in the chain, merged atoms are not
    z = np.concatenate(zz)
    zt = np.concatenate(zz_true)
    homogeneity = sum(np.bincount(zt[z == c]).max() for c in np.unique(z)) / len(z)
    assert homogeneity > 0.97, homogeneity


def test_kmax_exhaustion_is_an_error(ctx):
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 * k + 0.01 * rng.standard_normal(40) for k in range(6)])
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(x.mean()), 100.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=3)
    with pytest.raises(b.BiasError) as e:
        b.collapsed_gibbs_sampler(hdp, [x], [np.zeros(len(x), dtype=np.int64)], 5, 0, 1, ctx=ctx)
    assert e.value.code == 5
