"""ctypes front end of the CPU oracle (oracle/bias_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product path
(bias.jl_b200 -> libbias_cuda) never imports this module.

Parity status: conjugate arithmetic pinned against the reference's documented
REPL transcripts (tests/golden/reference_known_answers.json); sampler draws,
trajectories, the Stirling step and Distributions.jl draws are "parity unpinned"
(the reference has no tests and Julia 0.4 cannot run here).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbias_oracle.so")

MD, G1G1 = 0, 1
INT, SENT, F64 = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bias_oracle.c")
    hdr = os.path.join(_HERE, "bias_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libbias_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


class _Comps(C.Structure):
    _fields_ = [("kind", C.c_int32), ("K", C.c_int64), ("dd", C.c_int64), ("aa", C.c_double),
                ("mi", C.c_void_p), ("mm", C.c_void_p),
                ("m0", C.c_double), ("v0", C.c_double), ("vv", C.c_double),
                ("mu", C.c_void_p), ("nn", C.c_void_p)]


class _Data(C.Structure):
    _fields_ = [("kind", C.c_int32), ("w", C.c_void_p), ("x", C.c_void_p),
                ("sptr", C.c_void_p), ("sw", C.c_void_p), ("sc", C.c_void_p)]


class _Hdp(C.Structure):
    _fields_ = [("KK", C.c_int64), ("Kmax", C.c_int64),
                ("gg", C.c_double), ("g1", C.c_double), ("g2", C.c_double),
                ("aa", C.c_double), ("a1", C.c_double), ("a2", C.c_double),
                ("beta", C.c_void_p)]


class _Rcrp(C.Structure):
    _fields_ = [("KK", C.c_int64), ("Kmax", C.c_int64), ("TT", C.c_int64), ("aa", C.c_void_p),
                ("a1", C.c_double), ("a2", C.c_double)]


class _Crf(C.Structure):
    _fields_ = [("G", C.c_int64), ("ptr", C.c_void_p), ("tji", C.c_void_p), ("njt", C.c_void_p), ("kjt", C.c_void_p),
                ("n_tbls", C.c_void_p), ("mm", C.c_void_p)]


class _Rcrf(C.Structure):
    _fields_ = [("KK", C.c_int64), ("Kmax", C.c_int64), ("TT", C.c_int64), ("eptr", C.c_void_p), ("ptr", C.c_void_p),
                ("tji", C.c_void_p), ("njt", C.c_void_p), ("kjt", C.c_void_p), ("n_tbls", C.c_void_p), ("mm", C.c_void_p),
                ("gg", C.c_void_p), ("aa", C.c_void_p), ("g1", C.c_double), ("g2", C.c_double), ("a1", C.c_double),
                ("a2", C.c_double)]


class _Rng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_void_p
        i64, f64, i32 = C.c_int64, C.c_double, C.c_int
        L.orc_julia_sum.restype = f64
        L.orc_julia_sum.argtypes = [vp, i64]
        L.orc_lognormalize.argtypes = [vp, i64]
        L.orc_sample.restype = i64
        L.orc_sample.argtypes = [vp, i64, f64]
        L.orc_gen_bars.argtypes = [i64, i64, f64, vp]
        L.orc_gaussian1d_pdf.restype = f64
        L.orc_gaussian1d_pdf.argtypes = [f64, f64, f64]
        L.orc_gaussian1d_logpdf.restype = f64
        L.orc_gaussian1d_logpdf.argtypes = [f64, f64, f64]
        cp, dpp = C.POINTER(_Comps), C.POINTER(_Data)
        L.orc_additem.argtypes = [cp, i64, dpp, i64]
        L.orc_delitem.restype = i32
        L.orc_delitem.argtypes = [cp, i64, dpp, i64]
        for f in (L.orc_logpredictive, L.orc_loglikelihood):
            f.restype = f64
            f.argtypes = [cp, i64, dpp, i64]
        L.orc_g1g1_posterior.argtypes = [cp, i64, dp, dp]
        L.orc_comps_reset.argtypes = [cp, i64]
        L.orc_lda_init.restype = f64
        L.orc_lda_init.argtypes = [cp, dpp, i64, vp, vp, vp, i32]
        L.orc_lda_sweep.restype = f64
        L.orc_lda_sweep.argtypes = [cp, dpp, i64, vp, vp, vp, f64, vp, vp, vp, i32]
        L.orc_lda_conditional.restype = i64
        L.orc_lda_conditional.argtypes = [cp, dpp, i64, i64, i64, vp, f64, f64, vp, vp]
        L.orc_lda_sweep_linear.restype = None
        L.orc_lda_sweep_linear.argtypes = [cp, dpp, i64, vp, vp, vp, f64, vp, vp, vp]
        L.orc_lda_sweep_follow.restype = i64
        L.orc_lda_sweep_follow.argtypes = [cp, dpp, i64, vp, vp, vp, f64, vp, vp, C.POINTER(f64)]
        L.orc_bmm_init.restype = f64
        L.orc_bmm_init.argtypes = [cp, dpp, i64, vp, vp]
        L.orc_bmm_sweep.restype = f64
        L.orc_bmm_sweep.argtypes = [cp, dpp, i64, vp, vp, f64, vp, vp]
        L.orc_bmm_conditional.restype = i64
        L.orc_bmm_conditional.argtypes = [cp, dpp, i64, i64, vp, f64, f64, vp, vp]
        rp, hp = C.POINTER(_Rng), C.POINTER(_Hdp)
        L.orc_rng_seed.argtypes = [rp, C.c_uint64]
        for f in (L.orc_rand, L.orc_randn):
            f.restype = f64
            f.argtypes = [rp]
        L.orc_rand_gamma.restype = f64
        L.orc_rand_gamma.argtypes = [rp, f64]
        L.orc_rand_beta.restype = f64
        L.orc_rand_beta.argtypes = [rp, f64, f64]
        L.orc_hdp_init.restype = f64
        L.orc_hdp_init.argtypes = [hp, cp, dpp, i64, vp, vp, vp]
        L.orc_hdp_sweep.restype = f64
        L.orc_hdp_sweep.argtypes = [hp, cp, dpp, i64, vp, vp, vp, rp, vp, i32]
        L.orc_hdp_conditional.restype = i64
        L.orc_hdp_conditional.argtypes = [hp, cp, dpp, i64, i64, i64, vp, f64, vp, vp]
        L.orc_hdp_table_conditional.restype = i64
        L.orc_hdp_table_conditional.argtypes = [vp, i64, f64, f64, vp]
        L.orc_hdp_resample_beta.restype = i32
        L.orc_hdp_resample_beta.argtypes = [hp, i64, vp, vp, vp, i64, i32, i32, rp, vp]
        L.orc_hdp_sample_hyperparam.argtypes = [hp, i64, vp, i64, rp]
        L.orc_stirling_table.argtypes = [i64, vp]
        rcp = C.POINTER(_Rcrp)
        L.orc_rcrp_init.argtypes = [rcp, cp, dpp, vp, vp, vp]
        L.orc_rcrp_sweep.restype = f64
        L.orc_rcrp_sweep.argtypes = [rcp, cp, dpp, vp, vp, vp, rp, vp, i32, i32, i32]
        L.orc_rcrp_conditional.restype = i64
        L.orc_rcrp_conditional.argtypes = [rcp, cp, dpp, vp, i64, i64, i64, vp, f64, vp, vp]
        fp = C.POINTER(_Crf)
        L.orc_crf_init.argtypes = [hp, fp, cp, dpp, vp]
        L.orc_crf_sweep.restype = f64
        L.orc_crf_sweep.argtypes = [hp, fp, cp, dpp, vp, rp, vp, i32, i32, i32]
        L.orc_crf_table_conditional.restype = i64
        L.orc_crf_table_conditional.argtypes = [hp, fp, cp, dpp, i64, i64, vp, f64, vp, vp]
        rfp = C.POINTER(_Rcrf)
        L.orc_rcrf_init.argtypes = [rfp, cp, dpp, vp]
        L.orc_rcrf_sweep.restype = f64
        L.orc_rcrf_sweep.argtypes = [rfp, cp, dpp, vp, rp, vp, i32, i32, i32, i32]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------- common.jl
def julia_sum(a) -> float:
    a = _f64(a)
    return lib().orc_julia_sum(_p(a), a.size)


def lognormalize(pp) -> np.ndarray:
    pp = _f64(pp).copy()
    lib().orc_lognormalize(_p(pp), pp.size)
    return pp


def sample(w, r: float) -> int:
    w = _f64(w)
    return int(lib().orc_sample(_p(w), w.size, float(r)))


def gen_bars(n_bars: int, n_vocab: int, noise: float) -> np.ndarray:
    out = np.zeros((n_bars, n_vocab), dtype=np.float64)
    lib().orc_gen_bars(n_bars, n_vocab, float(noise), _p(out))
    return out


def gaussian1d_pdf(mu, vv, x):
    return lib().orc_gaussian1d_pdf(float(mu), float(vv), float(x))


def gaussian1d_logpdf(mu, vv, x):
    return lib().orc_gaussian1d_logpdf(float(mu), float(vv), float(x))


# ------------------------------------------------------------ conjugates.jl
class Comps:
    """K component slots of one conjugate family (struct of arrays)."""

    def __init__(self, kind: int, K: int, *, dd: int = 0, aa_total: float = 0.0,
                 m0: float = 0.0, v0: float = 1.0, vv: float = 1.0):
        self.kind, self.K = kind, K
        self.nn = np.zeros(K, dtype=np.int64)
        self.mi = np.zeros((K, max(dd, 1)), dtype=np.int64)
        self.mm = np.zeros(K, dtype=np.int64)
        self.mu = np.full(K, float(m0), dtype=np.float64)
        self.dd = dd
        # MultinomialDirichlet(dd, aa) stores aa/dd (src/conjugates.jl:142)
        self.aa = (aa_total / dd) if dd else 0.0
        self.m0, self.v0, self.vv = float(m0), float(v0), float(vv)
        self.c = _Comps(kind, K, dd, self.aa, _p(self.mi), _p(self.mm),
                        self.m0, self.v0, self.vv, _p(self.mu), _p(self.nn))

    @classmethod
    def md(cls, K, dd, aa_total):
        return cls(MD, K, dd=dd, aa_total=aa_total)

    @classmethod
    def g1g1(cls, K, m0, v0, vv):
        return cls(G1G1, K, m0=m0, v0=v0, vv=vv)

    def ref(self):
        return C.byref(self.c)

    def posterior(self, k):
        mu, vv = C.c_double(), C.c_double()
        lib().orc_g1g1_posterior(self.ref(), k, C.byref(mu), C.byref(vv))
        return mu.value, vv.value


class Data:
    def __init__(self, kind, *, w=None, x=None, sptr=None, sw=None, sc=None):
        self.kind = kind
        self.w = None if w is None else _i64(w)
        self.x = None if x is None else _f64(x)
        self.sptr = None if sptr is None else _i64(sptr)
        self.sw = None if sw is None else _i64(sw)
        self.sc = None if sc is None else _i64(sc)
        self.c = _Data(kind, _p(self.w), _p(self.x), _p(self.sptr), _p(self.sw), _p(self.sc))
        if kind == INT:
            self.n = self.w.size
        elif kind == F64:
            self.n = self.x.size
        else:
            self.n = self.sptr.size - 1

    @classmethod
    def words(cls, w):
        return cls(INT, w=w)

    @classmethod
    def reals(cls, x):
        return cls(F64, x=x)

    @classmethod
    def sents(cls, sptr, sw, sc):
        return cls(SENT, sptr=sptr, sw=sw, sc=sc)

    def ref(self):
        return C.byref(self.c)


def additem(c: Comps, k, d: Data, i):
    lib().orc_additem(c.ref(), k, d.ref(), i)


def delitem(c: Comps, k, d: Data, i) -> int:
    return lib().orc_delitem(c.ref(), k, d.ref(), i)


def logpredictive(c: Comps, k, d: Data, i) -> float:
    return lib().orc_logpredictive(c.ref(), k, d.ref(), i)


def loglikelihood(c: Comps, k, d: Data, i) -> float:
    return lib().orc_loglikelihood(c.ref(), k, d.ref(), i)


# ------------------------------------------------------------------- LDA.jl
class LDAState:
    """Serial LDA sampler state: components + nn[D,K] (src/LDA.jl:75-98)."""

    def __init__(self, comps: Comps, data: Data, ptr, z, alpha: float, with_diag: bool = True):
        self.c, self.d = comps, data
        self.ptr = _i64(ptr)
        self.D = self.ptr.size - 1
        self.z = _i64(z).copy()
        self.alpha = float(alpha)
        self.nn = np.zeros((self.D, comps.K), dtype=np.int64)
        self.loglik = lib().orc_lda_init(comps.ref(), data.ref(), self.D, _p(self.ptr), _p(self.z), _p(self.nn),
                                         int(with_diag))

    def sweep(self, uniforms, doc_order=None, tok_order=None, diag_mode=1) -> float:
        u = _f64(uniforms)
        do = None if doc_order is None else _i64(doc_order)
        to = None if tok_order is None else _i64(tok_order)
        self.loglik = lib().orc_lda_sweep(self.c.ref(), self.d.ref(), self.D, _p(self.ptr), _p(self.z),
                                          _p(self.nn), self.alpha, _p(do), _p(to), _p(u), diag_mode)
        return self.loglik

    def sweep_linear(self, uniforms, doc_order=None, tok_order=None):
        """The serial sweep with linear-domain Float64 weights (word ids; convergence fixtures)."""
        u = _f64(uniforms)
        do = None if doc_order is None else _i64(doc_order)
        to = None if tok_order is None else _i64(tok_order)
        lib().orc_lda_sweep_linear(self.c.ref(), self.d.ref(), self.D, _p(self.ptr), _p(self.z), _p(self.nn),
                                   self.alpha, _p(do), _p(to), _p(u))

    def sweep_follow(self, uniforms, other_z):
        """Serial sweep that adopts other_z wherever its own draw differs -> (number of disagreements, largest distance
        of a disagreeing uniform from the nearest edge of the oracle's CDF)."""
        u, oz = _f64(uniforms), _i64(other_z)
        worst = C.c_double(0.0)
        n = lib().orc_lda_sweep_follow(self.c.ref(), self.d.ref(), self.D, _p(self.ptr), _p(self.z), _p(self.nn),
                                       self.alpha, _p(u), _p(oz), C.byref(worst))
        return int(n), float(worst.value)

    def conditional(self, j, i, r):
        K = self.c.K
        raw, prob = np.zeros(K), np.zeros(K)
        k = lib().orc_lda_conditional(self.c.ref(), self.d.ref(), j, i, int(self.z[i]), _p(self.nn),
                                      self.alpha, float(r), _p(raw), _p(prob))
        return int(k), raw, prob


# ------------------------------------------------------------------- BMM.jl
class BMMState:
    def __init__(self, comps: Comps, data: Data, z, alpha: float):
        self.c, self.d = comps, data
        self.N = data.n
        self.z = _i64(z).copy()
        self.alpha = float(alpha)
        self.nn = np.zeros(comps.K, dtype=np.int64)
        self.loglik = lib().orc_bmm_init(comps.ref(), data.ref(), self.N, _p(self.z), _p(self.nn))

    def sweep(self, uniforms, order=None) -> float:
        u = _f64(uniforms)
        o = None if order is None else _i64(order)
        self.loglik = lib().orc_bmm_sweep(self.c.ref(), self.d.ref(), self.N, _p(self.z), _p(self.nn),
                                          self.alpha, _p(o), _p(u))
        return self.loglik

    def conditional(self, i, r):
        K = self.c.K
        raw, prob = np.zeros(K), np.zeros(K)
        k = lib().orc_bmm_conditional(self.c.ref(), self.d.ref(), i, int(self.z[i]), _p(self.nn),
                                      self.alpha, float(r), _p(raw), _p(prob))
        return int(k), raw, prob


# ------------------------------------------------------------------- HDP.jl
class Rng:
    def __init__(self, seed: int):
        self.r = _Rng()
        lib().orc_rng_seed(C.byref(self.r), seed)

    def ref(self):
        return C.byref(self.r)

    def rand(self):
        return lib().orc_rand(self.ref())

    def gamma(self, a):
        return lib().orc_rand_gamma(self.ref(), float(a))

    def beta(self, a, b):
        return lib().orc_rand_beta(self.ref(), float(a), float(b))


def stirling_table(nmax: int) -> np.ndarray:
    out = np.zeros(nmax * (nmax + 1) // 2, dtype=np.float64)
    lib().orc_stirling_table(nmax, _p(out))
    return out


def hdp_table_conditional(lstir, n, a, r):
    prob = np.zeros(n)
    m = lib().orc_hdp_table_conditional(_p(lstir), n, float(a), float(r), _p(prob))
    return int(m), prob


class HDPState:
    """Direct-assignment HDP state (src/HDP.jl:166-228). comps.K == Kmax; the last
    slot stays empty and stands for the prototype `hdp.component`."""

    def __init__(self, comps: Comps, data: Data, ptr, z, KK, gg, g1, g2, aa, a1, a2):
        self.c, self.d = comps, data
        self.ptr = _i64(ptr)
        self.D = self.ptr.size - 1
        self.z = _i64(z).copy()
        self.Kmax = comps.K
        self.beta = np.zeros(self.Kmax + 1, dtype=np.float64)
        self.h = _Hdp(KK, self.Kmax, gg, g1, g2, aa, a1, a2, _p(self.beta))
        self.nn = np.zeros((self.D, self.Kmax), dtype=np.int64)
        self.loglik = lib().orc_hdp_init(C.byref(self.h), comps.ref(), data.ref(), self.D, _p(self.ptr),
                                         _p(self.z), _p(self.nn))

    @property
    def KK(self):
        return int(self.h.KK)

    def sweep(self, rng: Rng, uniforms=None, shuffle=False) -> float:
        u = None if uniforms is None else _f64(uniforms)
        self.loglik = lib().orc_hdp_sweep(C.byref(self.h), self.c.ref(), self.d.ref(), self.D, _p(self.ptr),
                                          _p(self.z), _p(self.nn), rng.ref(), _p(u), int(shuffle))
        return self.loglik

    def conditional(self, j, i, r):
        n = self.KK + 1
        raw, prob = np.zeros(n), np.zeros(n)
        k = lib().orc_hdp_conditional(C.byref(self.h), self.c.ref(), self.d.ref(), j, i, int(self.z[i]),
                                      _p(self.nn), float(r), _p(raw), _p(prob))
        return int(k), raw, prob

    def resample_beta(self, lstir, nmax, n_internals, sample_hyper, rng: Rng):
        M = np.zeros((self.D, self.Kmax), dtype=np.int64)
        rc = lib().orc_hdp_resample_beta(C.byref(self.h), self.D, _p(self.ptr), _p(self.nn), _p(lstir), nmax,
                                         n_internals, int(sample_hyper), rng.ref(), _p(M))
        if rc != 0:
            raise IndexError("n_jk exceeds the Stirling table (the reference throws BoundsError)")
        return M


def hdp_sample_hyperparam(KK, gg, g1, g2, aa, a1, a2, ptr, m, rng: "Rng"):
    """One call of sample_hyperparam!(hdp, n_group_j, m) (src/HDP.jl:124-159) on a bare parameter set -> (aa, gg)."""
    beta = np.zeros(2)
    h = _Hdp(int(KK), int(KK) + 1, gg, g1, g2, aa, a1, a2, _p(beta))
    p = _i64(ptr)
    lib().orc_hdp_sample_hyperparam(C.byref(h), p.size - 1, _p(p), int(m), rng.ref())
    return float(h.aa), float(h.gg)


# ------------------------------------------------------------------ RCRP.jl
class RCRPState:
    """Recurrent CRP state (src/RCRP.jl:77-143). comps.K == Kmax; the last slot stays empty (prototype)."""

    def __init__(self, comps: Comps, data: Data, ptr, z, KK, aa, a1, a2):
        self.c, self.d = comps, data
        self.ptr = _i64(ptr)
        self.TT = self.ptr.size - 1
        self.z = _i64(z).copy()
        self.Kmax = comps.K
        self.aa = _f64(np.full(self.TT, aa) if np.isscalar(aa) else aa).copy()
        self.h = _Rcrp(KK, self.Kmax, self.TT, _p(self.aa), a1, a2)
        self.nn = np.zeros((self.TT, self.Kmax), dtype=np.int64)
        lib().orc_rcrp_init(C.byref(self.h), comps.ref(), data.ref(), _p(self.ptr), _p(self.z), _p(self.nn))

    @property
    def KK(self):
        return int(self.h.KK)

    def sweep(self, rng: Rng, uniforms=None, shuffle=False, sample_hyper=False, n_internals=10) -> float:
        u = None if uniforms is None else _f64(uniforms)
        return lib().orc_rcrp_sweep(C.byref(self.h), self.c.ref(), self.d.ref(), _p(self.ptr), _p(self.z), _p(self.nn),
                                    rng.ref(), _p(u), int(shuffle), int(sample_hyper), int(n_internals))

    def conditional(self, t, i, r):
        n = self.KK + 1
        raw, prob = np.zeros(n), np.zeros(n)
        k = lib().orc_rcrp_conditional(C.byref(self.h), self.c.ref(), self.d.ref(), _p(self.ptr), t, i, int(self.z[i]),
                                       _p(self.nn), float(r), _p(raw), _p(prob))
        return int(k), raw, prob


# ------------------------------------------------------------------ HDP.jl CRF_gibbs_sampler!
class CRFState:
    """Chinese-restaurant-franchise state (src/HDP.jl:402-470). comps.K == Kmax; the last slot stays empty (prototype)."""

    def __init__(self, comps: Comps, data: Data, ptr, z, KK, gg, g1, g2, aa, a1, a2):
        self.c, self.d = comps, data
        self.ptr = _i64(ptr)
        self.G = self.ptr.size - 1
        N = int(self.ptr[-1])
        self.z = _i64(z).copy()
        self.Kmax = comps.K
        self.beta = np.zeros(self.Kmax + 1)
        self.h = _Hdp(KK, self.Kmax, gg, g1, g2, aa, a1, a2, _p(self.beta))
        self.tji, self.njt, self.kjt = (np.zeros(max(N, 1), dtype=np.int64) for _ in range(3))
        self.n_tbls = np.zeros(max(self.G, 1), dtype=np.int64)
        self.mm = np.zeros(self.Kmax, dtype=np.int64)
        self.f = _Crf(self.G, _p(self.ptr), _p(self.tji), _p(self.njt), _p(self.kjt), _p(self.n_tbls), _p(self.mm))
        lib().orc_crf_init(C.byref(self.h), C.byref(self.f), comps.ref(), data.ref(), _p(self.z))

    @property
    def KK(self):
        return int(self.h.KK)

    def sweep(self, rng: Rng, uniforms=None, defer_deaths=False, sample_hyper=False, n_internals=10) -> float:
        u = None if uniforms is None else _f64(uniforms)
        return lib().orc_crf_sweep(C.byref(self.h), C.byref(self.f), self.c.ref(), self.d.ref(), _p(self.z), rng.ref(),
                                   _p(u), int(defer_deaths), int(sample_hyper), int(n_internals))

    def table_conditional(self, j, i, r):
        n = int(self.n_tbls[j]) + 1
        raw, prob = np.zeros(n), np.zeros(n)
        t = lib().orc_crf_table_conditional(C.byref(self.h), C.byref(self.f), self.c.ref(), self.d.ref(), j, i,
                                            _p(self.z), float(r), _p(raw), _p(prob))
        return int(t), raw, prob


# ------------------------------------------------------------------ RCRF.jl
class RCRFState:
    """Recurrent Chinese restaurant franchise (src/RCRF.jl:122-209).  eptr[TT+1]: first restaurant of each epoch;
    ptr[G+1]: first customer of each restaurant.  comps.K == Kmax; the last slot stays empty (prototype)."""

    def __init__(self, comps: Comps, data: Data, eptr, ptr, z, KK, gg, g1, g2, aa, a1, a2):
        self.c, self.d = comps, data
        self.eptr, self.ptr = _i64(eptr), _i64(ptr)
        self.TT, self.G = self.eptr.size - 1, self.ptr.size - 1
        N = int(self.ptr[-1])
        self.z = _i64(z).copy()
        self.Kmax = comps.K
        self.gg = _f64(np.full(self.TT, gg) if np.isscalar(gg) else gg).copy()
        self.aa = _f64(np.full(self.TT, aa) if np.isscalar(aa) else aa).copy()
        self.tji, self.njt, self.kjt = (np.zeros(max(N, 1), dtype=np.int64) for _ in range(3))
        self.n_tbls = np.zeros(max(self.G, 1), dtype=np.int64)
        self.mm = np.zeros((self.TT, self.Kmax), dtype=np.int64)
        self.f = _Rcrf(KK, self.Kmax, self.TT, _p(self.eptr), _p(self.ptr), _p(self.tji), _p(self.njt), _p(self.kjt),
                       _p(self.n_tbls), _p(self.mm), _p(self.gg), _p(self.aa), g1, g2, a1, a2)
        lib().orc_rcrf_init(C.byref(self.f), comps.ref(), data.ref(), _p(self.z))

    @property
    def KK(self):
        return int(self.f.KK)

    def sweep(self, rng: Rng, uniforms=None, defer_deaths=False, join_tables=True, sample_hyper=False, n_internals=10) -> float:
        u = None if uniforms is None else _f64(uniforms)
        return lib().orc_rcrf_sweep(C.byref(self.f), self.c.ref(), self.d.ref(), _p(self.z), rng.ref(), _p(u),
                                    int(defer_deaths), int(join_tables), int(sample_hyper), int(n_internals))
