"""CPU checks of the oracle's restatement of the Chinese-restaurant-franchise sampler (reference src/HDP.jl:402-709).
The reference has no tests for it ("parity unpinned"); these pin the invariants its loop maintains
and the equivalence of the two death orders (immediate renumbering vs. once per iteration)."""
import numpy as np

from oracle import oracle as orc


def _case(seed=0, G=6, n=40, Kmax=40, KK=3):
    rng = np.random.default_rng(seed)
    lens = rng.integers(n // 2, n, G)
    lens[1] = 0
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    N = int(ptr[-1])
    atoms = rng.integers(0, 5, N)
    x = 3.0 * (atoms + 1) + 0.05 * rng.standard_normal(N)
    z = rng.integers(0, KK, N)
    st = orc.CRFState(orc.Comps.g1g1(Kmax, float(x.mean()), 10.0, 0.0025), orc.Data.reals(x), ptr, z, KK, 1.0, 0.1, 0.1,
                      1.0, 0.1, 0.1)
    return st, x, ptr, atoms


def check_crf_invariants(z, ptr, tji, njt, kjt, n_tbls, mm, KK, comp_nn=None):
    G = len(ptr) - 1
    tables_per_dish = np.zeros(KK, dtype=np.int64)
    for j in range(G):
        p0, n = int(ptr[j]), int(ptr[j + 1] - ptr[j])
        nt = int(n_tbls[j])
        assert (nt == 0) == (n == 0)
        t = tji[p0:p0 + n]
        assert ((t >= 0) & (t < nt)).all()
        assert np.array_equal(njt[p0:p0 + nt], np.bincount(t, minlength=nt)) and (njt[p0:p0 + nt] > 0).all()
        assert np.array_equal(kjt[p0:p0 + nt][t], z[p0:p0 + n])          # a customer eats the dish of its table
        tables_per_dish += np.bincount(kjt[p0:p0 + nt], minlength=KK)
    assert np.array_equal(mm[:KK], tables_per_dish) and (mm[:KK] > 0).all()
    assert ((z >= 0) & (z < KK)).all()
    if comp_nn is not None:
        assert np.array_equal(comp_nn[:KK], np.bincount(z, minlength=KK))


def test_crf_init_builds_first_appearance_tables():
    st, x, ptr, _ = _case(1)
    check_crf_invariants(st.z, ptr, st.tji, st.njt, st.kjt, st.n_tbls, st.mm, st.KK, st.c.nn)
    p0 = int(ptr[0])
    seen = []
    for k in st.z[p0:int(ptr[1])]:
        if k not in seen:
            seen.append(int(k))
    assert list(st.kjt[p0:p0 + int(st.n_tbls[0])]) == seen             # unique(zz[jj]) order, src/HDP.jl:446


def test_crf_sweeps_keep_the_seating_consistent_in_both_death_orders():
    for defer in (False, True):
        st, x, ptr, _ = _case(2)
        r = orc.Rng(3)
        for s in range(8):
            ll = st.sweep(r, defer_deaths=defer, sample_hyper=(s < 5), n_internals=3)
            assert not np.isnan(ll)
            check_crf_invariants(st.z, ptr, st.tji, st.njt, st.kjt, st.n_tbls, st.mm, st.KK, st.c.nn)
            assert st.h.gg > 0 and st.h.aa > 0
        for k in range(st.KK):
            assert abs(st.c.mu[k] - x[st.z == k].mean()) < 1e-9


def test_crf_recovers_the_atoms():
    st, x, ptr, atoms = _case(4, G=12, n=60, KK=1)
    r = orc.Rng(5)
    for _ in range(25):
        assert not np.isnan(st.sweep(r))
    homog = sum(np.bincount(atoms[st.z == c]).max() for c in np.unique(st.z)) / len(x)
    assert homog > 0.98 and len(np.unique(atoms)) <= st.KK <= len(np.unique(atoms)) + 3


def test_crf_table_conditional_is_a_distribution_and_leaves_state_untouched():
    st, x, ptr, _ = _case(6)
    snap = (st.z.copy(), st.njt.copy(), st.c.mu.copy(), st.c.nn.copy())
    for j, i in ((0, 3), (2, int(ptr[2]) + 1), (5, int(ptr[6]) - 1)):
        t, raw, prob = st.table_conditional(j, i, 0.37)
        assert abs(prob.sum() - 1.0) < 1e-12 and 0 <= t <= st.n_tbls[j]
        cdf = np.cumsum(prob)
        assert cdf[t] >= 0.37 and (t == 0 or cdf[t - 1] < 0.37)
    assert all(np.array_equal(a, b_) for a, b_ in zip(snap, (st.z, st.njt, st.c.mu, st.c.nn)))


def test_crf_with_empty_and_single_customer_restaurants():
    """xx[jj] may be empty in the reference (the loops over 1:length(xx[jj]) simply do not run): no tables there."""
    rng = np.random.default_rng(1)
    lens = [0, 1, 35, 0, 2, 17, 0, 1, 0]
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    N = int(ptr[-1])
    x = 3.0 * (rng.integers(0, 4, N) + 1) + 0.05 * rng.standard_normal(N)
    st = orc.CRFState(orc.Comps.g1g1(48, float(x.mean()), 10.0, 0.0025), orc.Data.reals(x), ptr, rng.integers(0, 2, N), 2,
                      1.0, 0.1, 0.1, 1.0, 0.1, 0.1)
    r = orc.Rng(3)
    for _ in range(5):
        assert not np.isnan(st.sweep(r, sample_hyper=True))
        check_crf_invariants(st.z, ptr, st.tji, st.njt, st.kjt, st.n_tbls, st.mm, st.KK, st.c.nn)
        assert list(st.n_tbls[[0, 1, 3, 6, 7, 8]]) == [0, 1, 0, 0, 1, 0]
