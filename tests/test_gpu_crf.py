"""GPU parity tests for the Chinese-restaurant-franchise sampler (reference src/HDP.jl:402-709, CRF_gibbs_sampler!):
exact agreement with the serial sampler for a single restaurant fed the same uniforms, the seating invariants over
sweeps with many restaurants, and recovery of the atoms."""
import numpy as np
import pytest

import bias_jl_b200 as b
from oracle import oracle as orc
from test_oracle_crf import check_crf_invariants

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0, seed=1357)
    yield c
    c.close()


def _data(seed, lens, n_atoms=5, sd=0.05):
    rng = np.random.default_rng(seed)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    N = int(ptr[-1])
    atoms = rng.integers(0, n_atoms, N)
    x = 3.0 * (atoms + 1) + sd * rng.standard_normal(N)
    return x, ptr, atoms, rng


@pytest.mark.parametrize("lens", [[120], [1], [33]])
def test_single_restaurant_iteration_is_reproduced_exactly(ctx, lens):
    """One restaurant is processed by one warp in the reference's order, so with the same uniforms every draw of the
    iteration — tables, dishes of new tables, dishes of tables — must equal the serial sampler's (death order: once
    per iteration on both sides)."""
    x, ptr, atoms, rng = _data(3, lens)
    N, KK, Kmax = len(x), 3, 48
    z0 = rng.integers(0, KK, N).astype(np.int32)
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(x.mean()), 10.0, 0.0025), KK, 1.3, 0.1, 0.1, 0.9, 0.1, 0.1, Kmax=Kmax)
    data = b.FlatData.from_csr(b._lib.F64, N, group_ptr=ptr, x=x)
    rs = b.ResidentSampler(ctx, hdp, data, z0, sample_hyperparam=False, crf=True)
    st = orc.CRFState(orc.Comps.g1g1(Kmax, float(x.mean()), 10.0, 0.0025), orc.Data.reals(x), ptr, z0, KK, 1.3, 0.1, 0.1,
                      0.9, 0.1, 0.1)
    t0 = rs.crf_tables()
    assert np.array_equal(t0["tji"], st.tji[:N]) and np.array_equal(t0["mm"], st.mm[:KK])
    r = orc.Rng(1)
    for sweep in range(6):
        u = rng.random(3 * N)
        st.sweep(r, uniforms=u, defer_deaths=True)
        rs.crf_sweep_with_uniforms(u)
        z = rs.get_zz()
        t = rs.crf_tables()
        assert rs.current_KK() == st.KK, sweep
        assert np.array_equal(z, st.z), sweep
        nt = int(st.n_tbls[0])
        assert t["n_tbls"][0] == nt and np.array_equal(t["tji"], st.tji[:N])
        assert np.array_equal(t["njt"][:nt], st.njt[:nt]) and np.array_equal(t["kjt"][:nt], st.kjt[:nt])
        assert np.array_equal(t["mm"], st.mm[:st.KK])
    assert len(np.unique(st.z)) >= 1
    rs.close()


def test_many_restaurants_keep_the_seating_consistent(ctx):
    lens = np.random.default_rng(0).integers(20, 70, 200)
    lens[[3, 77]] = 0
    x, ptr, atoms, rng = _data(5, lens)
    N = len(x)
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(x.mean()), 10.0, 0.0025), 2, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=64)
    data = b.FlatData.from_csr(b._lib.F64, N, group_ptr=ptr, x=x)
    rs = b.ResidentSampler(ctx, hdp, data, rng.integers(0, 2, N).astype(np.int32), sample_hyperparam=True, n_internals=4,
                           crf=True)
    for sweep in range(12):
        rs.sweep(1)
        hp, _ = rs.hdp_state()
        z = rs.get_zz()
        t = rs.crf_tables()
        comp = rs.components()
        check_crf_invariants(z, ptr, t["tji"], t["njt"], t["kjt"], t["n_tbls"], t["mm"], hp.KK, comp["nn"])
        assert hp.gg > 0 and hp.aa > 0
        for k in range(hp.KK):
            assert abs(comp["mu"][k] - x[z == k].mean()) < 1e-9
    homog = sum(np.bincount(atoms[z == c]).max() for c in np.unique(z)) / N
    assert homog > 0.98 and 5 <= hp.KK <= 12, (homog, hp.KK)
    assert np.isfinite(rs.stats().log_likelihood)
    rs.close()


def test_crf_gibbs_sampler_entry_point(ctx):
    """demo_CRF_Gaussian1DGaussian1D.jl at reduced size: CRF-generated groups, atoms N(2k, 0.001), KK starts at 1."""
    zz_true, n_atoms = b.data.gen_crf_data([100] * 30, 1.0, 0.5, seed=3)
    rng = np.random.default_rng(4)
    xx = [2.0 * (z + 1) + np.sqrt(0.001) * rng.standard_normal(len(z)) for z in zz_true]
    allx = np.concatenate(xx)
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(allx.mean()), 10.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=64)
    zz = [np.zeros(len(x), dtype=np.int64) for x in xx]
    KK_list, KK_dict = b.CRF_gibbs_sampler(hdp, xx, zz, 15, 1, 8, True, 5, ctx=ctx)
    assert len(KK_list) == 8 and hdp.KK == KK_list[-1] and hdp.KK in KK_dict
    used = len(np.unique(np.concatenate(zz_true)))
    z, zt = np.concatenate(zz), np.concatenate(zz_true)
    homog = sum(np.bincount(zt[z == c]).max() for c in np.unique(z)) / len(z)
    assert homog > 0.97 and used - 1 <= hdp.KK <= used + 5, (homog, hdp.KK, used)
    # the serial sampler on the same data
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in xx])])
    st = orc.CRFState(orc.Comps.g1g1(64, float(allx.mean()), 10.0, 0.001), orc.Data.reals(allx), ptr,
                      np.zeros(len(allx), dtype=np.int64), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1)
    r = orc.Rng(5)
    for _ in range(20):
        assert not np.isnan(st.sweep(r, sample_hyper=True, n_internals=5))
    assert abs(hdp.KK - st.KK) <= 3, (hdp.KK, st.KK)


def _md_case(kind, seed, lens, V=25):
    rng = np.random.default_rng(seed)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    N = int(ptr[-1])
    if kind == "int":
        topics = b.data.gen_bars(6, V, 0.0)
        zt = rng.integers(0, 6, N)
        w = np.array([rng.choice(V, p=topics[k]) for k in zt])
        return b.MultinomialDirichlet(V, 2.5), orc.Data.words(w), lambda K: orc.Comps.md(K, V, 2.5), ptr, \
            b.FlatData.from_csr(b._lib.INT, N, group_ptr=ptr, word=w), N
    sents, _, _ = b.data.bars_bmm_sentences(N, 10, 6, V, seed)
    sp = np.concatenate([[0], np.cumsum([len(s_) for s_ in sents])]).astype(np.int64)
    sw, sc = np.concatenate([s_.w for s_ in sents]), np.concatenate([s_.c for s_ in sents])
    return b.MultinomialDirichlet(V, 2.5), orc.Data.sents(sp, sw, sc), lambda K: orc.Comps.md(K, V, 2.5), ptr, \
        b.FlatData.from_csr(b._lib.SENT, N, group_ptr=ptr, sent_ptr=sp, sent_word=sw, sent_count=sc), N


@pytest.mark.parametrize("kind", ["int", "sent"])
def test_single_restaurant_iteration_is_reproduced_exactly_for_multinomial_items(ctx, kind):
    """The same exact-reproduction check for word-id and Sent items (MultinomialDirichlet): a table's score for a dish is
    the sum of its customers' individual predictives, counts move with the whole table."""
    comp, odata, ocomp, ptr, data, N = _md_case(kind, 4, [70])
    KK, Kmax = 3, 40
    rng = np.random.default_rng(9)
    z0 = rng.integers(0, KK, N).astype(np.int32)
    hdp = b.HDP(comp, KK, 1.3, 0.1, 0.1, 0.9, 0.1, 0.1, Kmax=Kmax)
    rs = b.ResidentSampler(ctx, hdp, data, z0, sample_hyperparam=False, crf=True)
    st = orc.CRFState(ocomp(Kmax), odata, ptr, z0, KK, 1.3, 0.1, 0.1, 0.9, 0.1, 0.1)
    r = orc.Rng(1)
    for sweep in range(5):
        u = rng.random(3 * N)
        st.sweep(r, uniforms=u, defer_deaths=True)
        rs.crf_sweep_with_uniforms(u)
        t = rs.crf_tables()
        nt = int(st.n_tbls[0])
        assert rs.current_KK() == st.KK and np.array_equal(rs.get_zz(), st.z), sweep
        assert t["n_tbls"][0] == nt and np.array_equal(t["tji"], st.tji[:N]) and np.array_equal(t["kjt"][:nt], st.kjt[:nt])
        assert np.array_equal(t["mm"], st.mm[:st.KK])
        comp_g = rs.components()
        assert np.array_equal(comp_g["mi"], st.c.mi[:st.KK]) and np.array_equal(comp_g["mm"], st.c.mm[:st.KK])
    rs.close()


@pytest.mark.parametrize("kind", ["int", "sent"])
def test_many_restaurants_with_multinomial_items_keep_the_seating_consistent(ctx, kind):
    lens = np.random.default_rng(1).integers(15, 40, 60)
    comp, odata, ocomp, ptr, data, N = _md_case(kind, 6, lens)
    hdp = b.HDP(comp, 2, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=64)
    rs = b.ResidentSampler(ctx, hdp, data, np.random.default_rng(2).integers(0, 2, N).astype(np.int32), sample_hyperparam=True,
                           n_internals=3, crf=True)
    for sweep in range(8):
        rs.sweep(1)
        hp, _ = rs.hdp_state()
        z = rs.get_zz()
        t = rs.crf_tables()
        comp_g = rs.components()
        check_crf_invariants(z, ptr, t["tji"], t["njt"], t["kjt"], t["n_tbls"], t["mm"], hp.KK, comp_g["nn"])
    V = comp.dd
    mi = np.zeros((hp.KK, V), dtype=np.int64)
    if kind == "int":
        np.add.at(mi, (z, np.asarray(odata.w)), 1)
    else:
        item = np.repeat(np.arange(N), np.diff(odata.sptr))
        np.add.at(mi, (z[item], odata.sw), odata.sc)
    assert np.array_equal(comp_g["mi"], mi)
    rs.close()


def test_crf_argument_errors(ctx):
    rng = np.random.default_rng(0)
    x = [np.concatenate([10.0 * k + 0.01 * rng.standard_normal(20) for k in range(6)])]
    hdp = b.HDP(b.Gaussian1DGaussian1D(25.0, 100.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=3)
    with pytest.raises(b.BiasError) as e:
        b.CRF_gibbs_sampler(hdp, x, [np.zeros(120, dtype=np.int64)], 4, 0, 1, ctx=ctx)
    assert e.value.code == 5


@pytest.mark.parametrize("kind", ["f64", "int", "sent"])
def test_empty_and_single_customer_restaurants(ctx, kind):
    """Ragged input as the reference accepts it (xx[jj] may be empty): restaurants without customers have no tables,
    a single customer always has exactly one, and the seating of the others is unaffected."""
    lens = [0, 1, 35, 0, 2, 17, 0, 1, 0]
    if kind == "f64":
        x, ptr, _, _ = _data(8, lens)
        N = len(x)
        comp = b.Gaussian1DGaussian1D(float(x.mean()), 10.0, 0.0025)
        data = b.FlatData.from_csr(b._lib.F64, N, group_ptr=ptr, x=x)
    else:
        comp, _, _, ptr, data, N = _md_case(kind, 8, lens)
    hdp = b.HDP(comp, 2, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=48)
    rs = b.ResidentSampler(ctx, hdp, data, np.random.default_rng(3).integers(0, 2, N).astype(np.int32), sample_hyperparam=True,
                           n_internals=2, crf=True)
    for _ in range(5):
        rs.sweep(1)
        hp, _ = rs.hdp_state()
        t = rs.crf_tables()
        check_crf_invariants(rs.get_zz(), ptr, t["tji"], t["njt"], t["kjt"], t["n_tbls"], t["mm"], hp.KK, rs.components()["nn"])
        assert list(t["n_tbls"][[0, 1, 3, 6, 7, 8]]) == [0, 1, 0, 0, 1, 0]
    rs.close()


def test_crf_without_items(ctx):
    hdp = b.HDP(b.Gaussian1DGaussian1D(0.0, 10.0, 0.0025), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=8)
    data = b.FlatData.from_csr(b._lib.F64, 0, group_ptr=np.zeros(3, dtype=np.int64), x=np.zeros(0))
    rs = b.ResidentSampler(ctx, hdp, data, np.zeros(0, dtype=np.int32), sample_hyperparam=False, crf=True)
    rs.sweep(2)
    assert len(rs.get_zz()) == 0 and list(rs.crf_tables()["n_tbls"]) == [0, 0]
    rs.close()
