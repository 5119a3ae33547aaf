"""CPU checks of the oracle's restatement of the recurrent Chinese restaurant franchise (reference src/RCRF.jl:122-938;
no reference tests exist, "parity unpinned"): the seating invariants of every epoch, both death orders, joined tables,
and recovery of atoms that appear and disappear over the epochs."""
import numpy as np

from oracle import oracle as orc


def _case(seed=0, TT=4, G_per=4, n=30, Kmax=48, KK=3):
    rng = np.random.default_rng(seed)
    lens = rng.integers(n // 2, n, TT * G_per)
    lens[2] = 0
    eptr = np.arange(TT + 1, dtype=np.int64) * G_per
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    N = int(ptr[-1])
    epoch = np.repeat(np.arange(TT * G_per) // G_per, lens)
    atoms = (rng.integers(0, 3, N) + epoch // 2) % 5          # the set of atoms drifts over the epochs
    x = 3.0 * (atoms + 1) + 0.05 * rng.standard_normal(N)
    z = rng.integers(0, KK, N)
    st = orc.RCRFState(orc.Comps.g1g1(Kmax, float(x.mean()), 10.0, 0.0025), orc.Data.reals(x), eptr, ptr, z, KK,
                       1.0, 0.1, 0.1, 1.0, 0.1, 0.1)
    return st, x, eptr, ptr, atoms


def check_rcrf_invariants(z, eptr, ptr, tji, njt, kjt, n_tbls, mm, KK, comp_nn=None, joined=False):
    TT = len(eptr) - 1
    assert ((z >= 0) & (z < KK)).all()
    for t in range(TT):
        per_dish = np.zeros(KK, dtype=np.int64)
        for j in range(int(eptr[t]), int(eptr[t + 1])):
            p0, n, nt = int(ptr[j]), int(ptr[j + 1] - ptr[j]), int(n_tbls[j])
            assert (nt == 0) == (n == 0)
            tt_ = tji[p0:p0 + n]
            assert ((tt_ >= 0) & (tt_ < nt)).all()
            assert np.array_equal(njt[p0:p0 + nt], np.bincount(tt_, minlength=nt)) and (njt[p0:p0 + nt] > 0).all()
            assert np.array_equal(kjt[p0:p0 + nt][tt_], z[p0:p0 + n])
            per_dish += np.bincount(kjt[p0:p0 + nt], minlength=KK)
        assert np.array_equal(mm[t, :KK], per_dish), t
    assert (mm[:, :KK].sum(0) > 0).all() and (mm[:, KK:] == 0).all()
    if comp_nn is not None:
        assert np.array_equal(comp_nn[:KK], np.bincount(z, minlength=KK))


def test_rcrf_sweeps_keep_every_epoch_consistent():
    for defer in (False, True):
        for join in (True, False):
            st, x, eptr, ptr, _ = _case(1)
            check_rcrf_invariants(st.z, eptr, ptr, st.tji, st.njt, st.kjt, st.n_tbls, st.mm, st.KK, st.c.nn)
            r = orc.Rng(2)
            for s in range(6):
                assert not np.isnan(st.sweep(r, defer_deaths=defer, join_tables=join, sample_hyper=(s < 4), n_internals=3))
                check_rcrf_invariants(st.z, eptr, ptr, st.tji, st.njt, st.kjt, st.n_tbls, st.mm, st.KK, st.c.nn)
                assert (st.gg > 0).all() and (st.aa > 0).all()
            for k in range(st.KK):
                assert abs(st.c.mu[k] - x[st.z == k].mean()) < 1e-9


def test_rcrf_recovers_drifting_atoms():
    st, x, eptr, ptr, atoms = _case(3, TT=5, G_per=6, n=50, KK=1)
    r = orc.Rng(4)
    for _ in range(25):
        assert not np.isnan(st.sweep(r))
    homog = sum(np.bincount(atoms[st.z == c]).max() for c in np.unique(st.z)) / len(x)
    assert homog > 0.98 and len(np.unique(atoms)) <= st.KK <= len(np.unique(atoms)) + 4


def test_rcrf_with_empty_epochs_and_restaurants():
    rng = np.random.default_rng(2)
    eptr = np.array([0, 3, 3, 7, 8, 8], dtype=np.int64)                 # epochs 1 and 4 have no restaurants
    lens = [20, 0, 1, 0, 25, 30, 2, 12]
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    N = int(ptr[-1])
    x = 3.0 * (rng.integers(0, 4, N) + 1) + 0.05 * rng.standard_normal(N)
    st = orc.RCRFState(orc.Comps.g1g1(48, float(x.mean()), 10.0, 0.0025), orc.Data.reals(x), eptr, ptr, rng.integers(0, 2, N),
                       2, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1)
    r = orc.Rng(3)
    for _ in range(5):
        assert not np.isnan(st.sweep(r, sample_hyper=True, n_internals=2))
        check_rcrf_invariants(st.z, eptr, ptr, st.tji, st.njt, st.kjt, st.n_tbls, st.mm, st.KK, st.c.nn)
        assert (st.gg > 0).all() and (st.aa > 0).all() and not st.mm[1].any() and not st.mm[4].any()
