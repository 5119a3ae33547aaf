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
    assert homog > 0.98 and 5 <= hp.KK <= 9, (homog, hp.KK)
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
    assert homog > 0.97 and used - 1 <= hdp.KK <= used + 3, (homog, hdp.KK, used)
    # the serial sampler on the same data
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in xx])])
    st = orc.CRFState(orc.Comps.g1g1(64, float(allx.mean()), 10.0, 0.001), orc.Data.reals(allx), ptr,
                      np.zeros(len(allx), dtype=np.int64), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1)
    r = orc.Rng(5)
    for _ in range(20):
        assert not np.isnan(st.sweep(r, sample_hyper=True, n_internals=5))
    assert abs(hdp.KK - st.KK) <= 3, (hdp.KK, st.KK)


def test_crf_argument_errors(ctx):
    lda_like = b.HDP(b.MultinomialDirichlet(5, 1.0), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1)
    with pytest.raises(ValueError):
        b.CRF_gibbs_sampler(lda_like, [np.array([0, 1, 2])], [np.zeros(3, dtype=np.int64)], 1, 0, 1, ctx=ctx)
    rng = np.random.default_rng(0)
    x = [np.concatenate([10.0 * k + 0.01 * rng.standard_normal(20) for k in range(6)])]
    hdp = b.HDP(b.Gaussian1DGaussian1D(25.0, 100.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=3)
    with pytest.raises(b.BiasError) as e:
        b.CRF_gibbs_sampler(hdp, x, [np.zeros(120, dtype=np.int64)], 4, 0, 1, ctx=ctx)
    assert e.value.code == 5
