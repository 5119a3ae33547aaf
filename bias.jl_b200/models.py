"""Host-side mirror of the reference's sampler API for the assignment-sweep hot path.

  LDA(c, KK, aa)                      src/LDA.jl:11-18
  BMM(c, KK, aa)                      src/BMM.jl:11-18
  HDP(c, KK, gg, g1, g2, aa, a1, a2)  src/HDP.jl:15-33
  init_zz(model, zz)                  src/common.jl:25-34
  collapsed_gibbs_sampler(model, xx, zz, n_burnins, n_lags, n_samples, ...)
                                      src/LDA.jl:67-148, src/BMM.jl:46-130, src/HDP.jl:166-399
  posterior(model, xx, zz)            src/LDA.jl:151-167, src/BMM.jl:134-151
  CRF_gibbs_sampler(hdp, xx, zz, n_burnins, n_lags, n_samples, ...)   src/HDP.jl:402-709
  RCRP(c, KK, aa, TT, a1, a2)         src/RCRP.jl:18-28
  RCRF(c, KK, gg, g1, g2, aa, a1, a2, TT), RCRF_gibbs_sampler(rcrf, xx, zz, ...)   src/RCRF.jl:16-37, 122-938
  RCRP_gibbs_sampler(rcrp, xx, zz, n_burnins, n_lags, n_samples, ...)   src/RCRP.jl:77-416
  posterior(rcrp, xx, KK_dict, KK)    src/RCRP.jl:418-453

Everything below the sampler entry point runs in libbias_cuda on the GPU; this module only
flattens the ragged Julia-style containers to CSR, calls the C ABI and writes zz back in place.
Python indices are 0-based.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import Context, check, lib, ptr
from .conjugates import Gaussian1DGaussian1D, MultinomialDirichlet, Sent


# ------------------------------------------------------------------ model types
class LDA:
    """Latent Dirichlet Allocation: finite KK, grouped data (src/LDA.jl:11-18)."""
    model_id = _lib.MODEL_LDA

    def __init__(self, component, KK: int, aa: float):
        self.component, self.KK, self.aa = component, int(KK), float(aa)

    def __repr__(self):
        return f"Latent Dirichlet Allocation model with {self.KK} {type(self.component).__name__} components"


class BMM:
    """Bayesian (finite) mixture model over flat data (src/BMM.jl:11-18)."""
    model_id = _lib.MODEL_BMM

    def __init__(self, component, KK: int, aa: float):
        self.component, self.KK, self.aa = component, int(KK), float(aa)

    def __repr__(self):
        return f"Finite Mixture Model with {self.KK} {type(self.component).__name__} components"


class HDP:
    """Hierarchical Dirichlet process mixture (src/HDP.jl:15-33). KK, gg and aa are updated in
    place by the sampler, as on the Julia object.  Kmax is the engine's slot capacity."""
    model_id = _lib.MODEL_HDP

    def __init__(self, component, KK: int, gg: float, g1: float, g2: float, aa: float, a1: float, a2: float,
                 Kmax: int = 256):
        self.component, self.KK = component, int(KK)
        self.gg, self.g1, self.g2 = float(gg), float(g1), float(g2)
        self.aa, self.a1, self.a2 = float(aa), float(a1), float(a2)
        self.Kmax = int(Kmax)

    def __repr__(self):
        return (f"Hierarchical Dirichlet Process Mixture Model with {self.KK} "
                f"{type(self.component).__name__} components")


class RCRP:
    """Recurrent Chinese restaurant process over TT epochs (src/RCRP.jl:18-28).  aa holds one concentration
    per epoch; KK and aa are updated in place by the sampler.  Kmax is the engine's slot capacity."""
    model_id = _lib.MODEL_RCRP

    def __init__(self, component, KK: int, aa, TT: int | None = None, a1: float = 1.0, a2: float = 1.0,
                 Kmax: int = 256):
        self.component, self.KK = component, int(KK)
        if np.isscalar(aa):
            if TT is None:
                raise ValueError("RCRP(c, KK, aa, TT, a1, a2): TT is required when aa is a scalar")
            self.aa = np.full(int(TT), float(aa))
        else:
            self.aa = np.array(aa, dtype=np.float64)
        self.a1, self.a2, self.Kmax = float(a1), float(a2), int(Kmax)

    def __repr__(self):
        return (f"Recurrent Chinese Restaurant Process with {self.KK} "
                f"{type(self.component).__name__} components")


class RCRF:
    """Recurrent Chinese restaurant franchise over TT epochs (src/RCRF.jl:16-37).  gg and aa hold one concentration per
    epoch; KK, gg and aa are updated in place by the sampler.  Kmax is the engine's dish-slot capacity."""
    model_id = _lib.MODEL_RCRF

    def __init__(self, component, KK: int, gg, g1: float, g2: float, aa, a1: float, a2: float, TT: int | None = None,
                 Kmax: int = 256):
        self.component, self.KK = component, int(KK)
        if (np.isscalar(gg) or np.isscalar(aa)) and TT is None:
            raise ValueError("RCRF(c, KK, gg, g1, g2, aa, a1, a2, TT): TT is required when gg / aa are scalars")
        self.gg = np.full(int(TT), float(gg)) if np.isscalar(gg) else np.array(gg, dtype=np.float64)
        self.aa = np.full(int(TT), float(aa)) if np.isscalar(aa) else np.array(aa, dtype=np.float64)
        self.g1, self.g2, self.a1, self.a2, self.Kmax = float(g1), float(g2), float(a1), float(a2), int(Kmax)

    def __repr__(self):
        return (f"Recurrent Chinese Restaurant Franchise Mixture Model with {self.KK} "
                f"{type(self.component).__name__} components")


def init_zz(model, zz, rng=None):
    """src/common.jl:25-34 — uniform random initial assignments (0-based)."""
    rng = rng or np.random.default_rng()
    if isinstance(model, BMM):
        zz[:] = rng.integers(0, model.KK, len(zz))
    else:
        for jj in range(len(zz)):
            zz[jj] = rng.integers(0, model.KK, len(zz[jj])).astype(np.int64)
    return zz


# ------------------------------------------------------------------ flattening
class FlatData:
    """CSR view of the reference's containers: xx::Vector{T} or Vector{Vector{T}}."""

    def __init__(self, component, xx, grouped: bool):
        items = [x for g in xx for x in g] if grouped else list(xx)
        if grouped:
            lens = np.array([len(g) for g in xx], dtype=np.int64)
            self.group_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            self.n_groups = len(xx)
        else:
            self.group_ptr, self.n_groups = None, 0
        self.n_items = len(items)
        probe = items[0] if items else (0.0 if isinstance(component, Gaussian1DGaussian1D) else 0)
        self.datum = component.datum_kind(probe)
        self.word = self.x = self.sent_ptr = self.sent_word = self.sent_count = None
        if self.datum == _lib.INT:
            if grouped and all(isinstance(g, np.ndarray) for g in xx) and len(xx):
                self.word = _lib.as_i32(np.concatenate(xx)) if self.n_items else np.zeros(0, np.int32)
            else:
                self.word = _lib.as_i32(items)
        elif self.datum == _lib.F64:
            if grouped and all(isinstance(g, np.ndarray) for g in xx) and len(xx):
                self.x = _lib.as_f64(np.concatenate(xx)) if self.n_items else np.zeros(0)
            else:
                self.x = _lib.as_f64(items)
        else:
            # merge duplicate words inside a Sent: the reference accumulates them into one
            # vocabulary slot (lll[w] += c, src/conjugates.jl:186-191)
            ws, cs, sp = [], [], [0]
            for s in items:
                uw, inv = np.unique(s.w, return_inverse=True)
                uc = np.zeros(len(uw), dtype=np.int64)
                np.add.at(uc, inv, s.c)
                ws.append(uw)
                cs.append(uc)
                sp.append(sp[-1] + len(uw))
            self.sent_ptr = np.asarray(sp, dtype=np.int64)
            self.sent_word = _lib.as_i32(np.concatenate(ws)) if ws else np.zeros(0, np.int32)
            self.sent_count = _lib.as_i32(np.concatenate(cs)) if cs else np.zeros(0, np.int32)
        self.c = _lib.Data(self.n_items, ptr(self.word), ptr(self.x), ptr(self.sent_ptr), ptr(self.sent_word),
                           ptr(self.sent_count))

    @classmethod
    def from_csr(cls, datum: int, n_items: int, group_ptr=None, word=None, x=None, sent_ptr=None, sent_word=None,
                 sent_count=None):
        """Already-flat inputs (what the Julia shim and bench.py hand over)."""
        self = cls.__new__(cls)
        self.datum, self.n_items = datum, int(n_items)
        self.group_ptr = None if group_ptr is None else _lib.as_i64(group_ptr)
        self.n_groups = 0 if group_ptr is None else len(self.group_ptr) - 1
        self.word = None if word is None else _lib.as_i32(word)
        self.x = None if x is None else _lib.as_f64(x)
        self.sent_ptr = None if sent_ptr is None else _lib.as_i64(sent_ptr)
        self.sent_word = None if sent_word is None else _lib.as_i32(sent_word)
        self.sent_count = None if sent_count is None else _lib.as_i32(sent_count)
        self.c = _lib.Data(self.n_items, ptr(self.word), ptr(self.x), ptr(self.sent_ptr), ptr(self.sent_word),
                           ptr(self.sent_count))
        return self


def _flat_zz(zz, grouped: bool) -> np.ndarray:
    if grouped:
        return _lib.as_i32(np.concatenate([np.asarray(g) for g in zz])) if len(zz) else np.zeros(0, np.int32)
    return _lib.as_i32(zz)


def _scatter_zz(flat: np.ndarray, zz, grouped: bool, group_ptr):
    if grouped:
        for jj in range(len(zz)):
            seg = flat[group_ptr[jj]:group_ptr[jj + 1]]
            if isinstance(zz[jj], np.ndarray) and zz[jj].shape == seg.shape:
                zz[jj][:] = seg
            else:
                zz[jj] = seg.astype(np.int64)
    else:
        zz[:] = flat


def _hdp_params(hdp: HDP, sample_hyperparam=True, n_internals=10) -> _lib.HdpParams:
    return _lib.HdpParams(hdp.KK, hdp.Kmax, hdp.gg, hdp.g1, hdp.g2, hdp.aa, hdp.a1, hdp.a2,
                          int(bool(sample_hyperparam)), int(n_internals))


_default_ctx: Context | None = None


def default_context(seed: int = 0) -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0, seed)
    return _default_ctx


# ------------------------------------------------------------------ resident sampler state
class _BatchShape:
    """What ResidentSampler needs to know about a batch it did not receive as FlatData (reload_u16)."""

    def __init__(self, n_items: int, n_groups: int, group_ptr: np.ndarray):
        self.n_items, self.n_groups, self.group_ptr = n_items, n_groups, group_ptr


class ResidentSampler:
    """bias_model: xx, zz and the component statistics live in HBM between sweeps."""

    def __init__(self, ctx: Context, model, data: FlatData, zz_flat: np.ndarray, *, sample_hyperparam=True,
                 n_internals=10, crf: bool = False, epoch_ptr=None, join_tables: bool = True):
        self.ctx, self.model, self.data = ctx, model, data
        self.n_internals = int(n_internals)
        self.spec = model.component.spec(data.datum)
        self.h = C.c_void_p()
        hp = _hdp_params(model, sample_hyperparam, n_internals) if isinstance(model, HDP) else None
        self.zz = _lib.as_i32(zz_flat)
        if isinstance(model, RCRF):
            self.epoch_ptr = _lib.as_i64(epoch_ptr)
            rp = _lib.RcrfParams(model.KK, model.Kmax, model.g1, model.g2, model.a1, model.a2, int(bool(sample_hyperparam)),
                                 int(n_internals), int(bool(join_tables)))
            gg, aa = _lib.as_f64(model.gg), _lib.as_f64(model.aa)
            if len(gg) != len(self.epoch_ptr) - 1 or len(aa) != len(gg):
                raise ValueError(f"RCRF: gg / aa have {len(gg)} / {len(aa)} entries for {len(self.epoch_ptr) - 1} epochs")
            check(lib().bias_rcrf_create(ctx.h, C.byref(self.spec), C.byref(data.c), data.n_groups, ptr(data.group_ptr),
                                         len(gg), ptr(self.epoch_ptr), ptr(self.zz), C.byref(rp), ptr(gg), ptr(aa),
                                         C.byref(self.h)))
            return
        if isinstance(model, RCRP):
            rp = _lib.RcrpParams(model.KK, model.Kmax, model.a1, model.a2, int(bool(sample_hyperparam)),
                                 int(n_internals))
            aa = _lib.as_f64(model.aa)
            if len(aa) != data.n_groups:
                raise ValueError(f"RCRP: aa has {len(aa)} entries for {data.n_groups} epochs")
            check(lib().bias_rcrp_create(ctx.h, C.byref(self.spec), C.byref(data.c), data.n_groups,
                                         ptr(data.group_ptr), ptr(self.zz), C.byref(rp), ptr(aa), C.byref(self.h)))
            return
        self.crf = bool(crf)
        if crf and not isinstance(model, HDP):
            raise TypeError("the CRF sampler is a sampler of the HDP (src/HDP.jl:402)")
        check(lib().bias_model_create(ctx.h, _lib.MODEL_CRF if crf else model.model_id, C.byref(self.spec), C.byref(data.c), data.n_groups,
                                      ptr(data.group_ptr), ptr(self.zz), model.KK, model.aa,
                                      C.byref(hp) if hp is not None else None, C.byref(self.h)))

    def reload(self, data: FlatData, zz_flat: np.ndarray):
        """Re-feed the resident model with a new batch (LDA / BMM): no allocation, buffers are reused."""
        self.data = data
        self.zz = _lib.as_i32(zz_flat)
        check(lib().bias_model_reload(self.h, C.byref(data.c), data.n_groups, ptr(data.group_ptr), ptr(self.zz)))

    def reload_u16(self, words16: np.ndarray, group_ptr: np.ndarray, zz16: np.ndarray):
        """reload() with word ids and assignments as uint16 host arrays (LDA x word ids, dd and KK <= 65536): half
        the bytes over PCIe.  Afterwards `self.data` only records the shape of the batch (the library owns its copy)."""
        w, z, gp = np.ascontiguousarray(words16), np.ascontiguousarray(zz16), _lib.as_i64(group_ptr)
        assert w.dtype == np.uint16 and z.dtype == np.uint16 and w.size == z.size
        check(lib().bias_model_reload_u16(self.h, w.size, ptr(w), gp.size - 1, ptr(gp), ptr(z)))
        self.data = _BatchShape(int(w.size), int(gp.size - 1), gp)

    def get_zz_u16(self, out: np.ndarray) -> np.ndarray:
        assert out.dtype == np.uint16 and out.flags.c_contiguous
        check(lib().bias_model_get_zz_u16(self.h, ptr(out)))
        return out

    # -- sweeps
    def sweep(self, n: int = 1):
        check(lib().bias_model_sweep(self.h, n))

    def stats(self) -> _lib.SweepStats:
        s = _lib.SweepStats()
        check(lib().bias_model_last_stats(self.h, C.byref(s)))
        return s

    def last_sweep_ms(self):
        ms, n = C.c_float(), C.c_int32()
        check(lib().bias_model_last_sweep_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- state access
    def get_zz(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.data.n_items, dtype=np.int32)
        assert out.dtype == np.int32 and out.size == self.data.n_items and out.flags.c_contiguous
        check(lib().bias_model_get_zz(self.h, ptr(out)))
        return out

    def set_zz(self, zz_flat):
        z = _lib.as_i32(zz_flat)
        check(lib().bias_model_set_zz(self.h, ptr(z)))

    def current_KK(self) -> int:
        return int(self.stats().KK)

    def components(self):
        K = self.current_KK()
        nn = np.zeros(K, dtype=np.int64)
        if self.spec.kind == _lib.MD:
            mi = np.zeros((K, self.spec.dd), dtype=np.int64)
            mm = np.zeros(K, dtype=np.int64)
            check(lib().bias_model_get_components(self.h, ptr(mi), ptr(mm), None, ptr(nn)))
            return {"mi": mi, "mm": mm, "nn": nn}
        mu = np.zeros(K)
        check(lib().bias_model_get_components(self.h, None, None, ptr(mu), ptr(nn)))
        return {"mu": mu, "nn": nn}

    def group_counts(self) -> np.ndarray:
        K = self.current_KK()
        nn = np.zeros((self.data.n_groups, K), dtype=np.int64)
        check(lib().bias_model_get_group_counts(self.h, ptr(nn)))
        return nn

    def hdp_state(self):
        hp = _lib.HdpParams()
        beta = np.zeros(self.model.Kmax + 1)
        check(lib().bias_model_get_hdp(self.h, C.byref(hp), ptr(beta)))
        return hp, beta[:hp.KK + 1].copy()

    # -- CRF seating (src/HDP.jl:420-425)
    def crf_tables(self):
        """-> dict(tji, njt, kjt, n_tbls, mm): njt / kjt hold the tables of group j at group_ptr[j] .. + n_tbls[j]"""
        N, G, KK = self.data.n_items, self.data.n_groups, self.current_KK()
        mm_shape = (len(self.epoch_ptr) - 1, KK) if isinstance(self.model, RCRF) else (KK,)
        out = {"tji": np.zeros(N, np.int32), "njt": np.zeros(N, np.int32), "kjt": np.zeros(N, np.int32),
               "n_tbls": np.zeros(G, np.int32), "mm": np.zeros(mm_shape, np.int32)}
        check(lib().bias_crf_get_tables(self.h, ptr(out["tji"]), ptr(out["njt"]), ptr(out["kjt"]), ptr(out["n_tbls"]),
                                        ptr(out["mm"])))
        return out

    def rcrf_state(self):
        """-> (KK, gg[TT], aa[TT])"""
        TT = len(self.epoch_ptr) - 1
        kk, gg, aa = C.c_int32(), np.zeros(TT), np.zeros(TT)
        check(lib().bias_model_get_rcrf(self.h, C.byref(kk), ptr(gg), ptr(aa)))
        return kk.value, gg, aa

    def crf_sweep_with_uniforms(self, uniforms):
        u = _lib.as_f64(uniforms)
        assert u.size == 3 * self.data.n_items
        check(lib().bias_crf_sweep_with_uniforms(self.h, ptr(u)))

    def rcrp_state(self):
        rp = _lib.RcrpParams()
        aa = np.zeros(self.data.n_groups)
        check(lib().bias_model_get_rcrp(self.h, C.byref(rp), ptr(aa)))
        return rp, aa

    def set_sample_hyperparam(self, on: bool):
        check(lib().bias_model_set_sample_hyperparam(self.h, int(bool(on))))

    # -- verification
    def conditionals(self, item_idx, uniforms):
        idx, u = _lib.as_i64(item_idx), _lib.as_f64(uniforms)
        n_out = self.current_KK() + (1 if isinstance(self.model, (HDP, RCRP, RCRF)) else 0)
        raw, prob = np.zeros((len(idx), n_out)), np.zeros((len(idx), n_out))
        draw = np.zeros(len(idx), dtype=np.int32)
        check(lib().bias_model_conditionals(self.h, len(idx), ptr(idx), ptr(u), ptr(raw), ptr(prob), ptr(draw)))
        return raw, prob, draw

    def frozen_draws(self, uniforms) -> np.ndarray:
        u = _lib.as_f64(uniforms)
        assert len(u) == self.data.n_items
        out = np.zeros(self.data.n_items, dtype=np.int32)
        check(lib().bias_model_frozen_draws(self.h, ptr(u), ptr(out)))
        return out

    def frozen_draws_dev(self, uniforms_dev_ptr: int, draws_dev_ptr: int):
        """frozen_draws with both arrays in device memory (raw pointers: float64[n_items], int32[n_items])."""
        check(lib().bias_model_frozen_draws_dev(self.h, C.c_void_p(uniforms_dev_ptr), C.c_void_p(draws_dev_ptr)))

    def sweep_with_uniforms(self, uniforms):
        """One production sweep that draws with the caller's uniforms (verification; a one-document corpus is swept
        serially exactly as src/LDA.jl:116-131)."""
        u = _lib.as_f64(uniforms)
        assert u.size == self.data.n_items
        check(lib().bias_model_sweep_with_uniforms(self.h, ptr(u)))

    # -- sharded runs inside the library (csrc/comm.cu)
    def share_prepare(self, rank: int, world: int, n_items_cap_global: int):
        check(lib().bias_model_share_prepare(self.h, rank, world, n_items_cap_global))

    def share_export(self) -> bytes:
        h = _lib.ShareHandle()
        check(lib().bias_model_share_export(self.h, C.byref(h)))
        return bytes(h)

    def share_attach(self, handles: list[bytes]):
        arr = (_lib.ShareHandle * len(handles))(*[_lib.ShareHandle.from_buffer_copy(b) for b in handles])
        check(lib().bias_model_share_attach(self.h, arr, len(handles)))

    def narrow_info(self):
        """-> (sweeps run on the narrow tables, number of hot words, hot rows allocated)"""
        cur, n_hot, cap = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib().bias_model_narrow_info(self.h, C.byref(cur), C.byref(n_hot), C.byref(cap)))
        return bool(cur.value), n_hot.value, cap.value

    def last_exchange(self):
        """-> (ms of the most recent exchange kernel, bytes pulled from the peers in it)"""
        ms, nb = C.c_float(), C.c_int64()
        check(lib().bias_model_last_exchange(self.h, C.byref(ms), C.byref(nb)))
        return ms.value, nb.value

    def replica_digest(self) -> int:
        d = C.c_uint64()
        check(lib().bias_model_replica_digest(self.h, C.byref(d)))
        return d.value

    def log_likelihood(self):
        """(sum_i log p(x_i | point estimates), n_items); perplexity = exp(-sum / n)."""
        s, n = C.c_double(), C.c_int64()
        check(lib().bias_model_log_likelihood(self.h, C.byref(s), C.byref(n)))
        return s.value, n.value

    def heldout_log_likelihood(self, heldout):
        """Document completion: heldout[j] = held-out word ids of group j.  -> (sum log p, n_tokens);
        held-out perplexity = exp(-sum / n_tokens)."""
        lens = np.array([len(h) for h in heldout], dtype=np.int64)
        hp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        hw = _lib.as_i32(np.concatenate([np.asarray(h) for h in heldout])) if hp[-1] else np.zeros(0, np.int32)
        assert len(hp) == self.data.n_groups + 1
        s, n = C.c_double(), C.c_int64()
        check(lib().bias_model_heldout_log_likelihood(self.h, ptr(hp), ptr(hw), C.byref(s), C.byref(n)))
        return s.value, n.value

    # -- sharded HDP phases (see include/bias_cuda.h and sharding.ShardedHDP)
    def set_host_seed(self, seed: int):
        check(lib().bias_model_set_host_seed(self.h, C.c_uint64(seed & (2**64 - 1))))

    def hdp_shard_pass(self):
        check(lib().bias_hdp_shard_pass(self.h))

    def hdp_shard_births(self):
        """-> (KK, beta[Kmax+1]) after this rank's births"""
        kk = C.c_int32()
        beta = np.zeros(self.model.Kmax + 1)
        check(lib().bias_hdp_shard_births(self.h, C.byref(kk), ptr(beta)))
        return kk.value, beta

    def hdp_shard_adopt(self, KK: int, beta: np.ndarray):
        bb = _lib.as_f64(beta)
        check(lib().bias_hdp_shard_adopt(self.h, int(KK), ptr(bb)))

    def hdp_shard_prune(self):
        check(lib().bias_hdp_shard_prune(self.h))

    def hdp_shard_tables(self, internal: int):
        """-> [(device_ptr, n, 'int64'), (device_ptr, 2, 'float64')]: this rank's table counts and alpha0 auxiliaries"""
        pt, ph, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        check(lib().bias_hdp_shard_tables(self.h, internal, C.byref(pt), C.byref(n), C.byref(ph)))
        return [(pt.value, n.value, "int64"), (ph.value, 2, "float64")]

    def hdp_shard_draw(self):
        check(lib().bias_hdp_shard_draw(self.h))

    def hdp_shard_end(self):
        check(lib().bias_hdp_shard_end(self.h))

    # -- multi-GPU exchange (AD-LDA)
    def delta_begin(self):
        check(lib().bias_model_delta_begin(self.h))

    def delta_begin_zero(self):
        check(lib().bias_model_delta_begin_zero(self.h))

    def delta_pack(self):
        """-> [(device_ptr, n_elems, 'int32'), (device_ptr, n_elems, 'float64')] (second may be empty)"""
        p32, p64, n32, n64 = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        check(lib().bias_model_delta_pack(self.h, C.byref(p32), C.byref(n32), C.byref(p64), C.byref(n64)))
        out = [(p32.value, n32.value, "int32")]
        if n64.value:
            out.append((p64.value, n64.value, "float64"))
        return out

    def delta_apply(self):
        check(lib().bias_model_delta_apply(self.h))

    def delta_pack16(self, world: int):
        """-> [(ptr, n, 'int32'), (ptr, n, 'int32')] or None when the model has no 16-bit exchange (not MD counts).
        The first buffer holds two 16-bit word-topic deltas per int32 (d1 * 65536 + d0: NCCL has no int16 sum, an
        int32 sum of this encoding is exact while the summed deltas fit 16 bits)."""
        if self.spec.kind != _lib.MD:
            return None
        p16, ps, n16, ns = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        check(lib().bias_model_delta_pack16(self.h, int(world), C.byref(p16), C.byref(n16), C.byref(ps), C.byref(ns)))
        return [(p16.value, n16.value // 2, "int32"), (ps.value, ns.value, "int32")]

    def delta_apply16(self):
        check(lib().bias_model_delta_apply16(self.h))

    def close(self):
        if self.h:
            lib().bias_model_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------ the reference entry points
class Group:
    """bias_group: resident LDA shards of THIS process (one per GPU, or several on one GPU) swept together, with the
    library's own count exchange after every sweep.  The shards stay owned by the caller."""

    def __init__(self, samplers):
        self.samplers = list(samplers)
        self.h = C.c_void_p()
        arr = (C.c_void_p * len(self.samplers))(*[s.h for s in self.samplers])
        check(lib().bias_group_create(arr, len(self.samplers), C.byref(self.h)))

    def sweep(self, n: int = 1):
        check(lib().bias_group_sweep(self.h, n))

    def close(self):
        if self.h:
            lib().bias_group_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def lda_run_multi(devices, lda: LDA, xx, zz, n_sweeps: int, seed: int = 0):
    """bias_lda_run_multi: collapsed_gibbs_sampler!(lda, xx, zz, ...) over several GPUs of this process; zz is mutated in
    place like the reference's.  Returns the per-sweep stats."""
    data = FlatData(lda.component, xx, True)
    if data.datum != _lib.INT:
        raise TypeError("lda_run_multi covers word-id observations")
    flat = _lib.as_i32(_flat_zz(zz, True))
    dev = _lib.as_i32(devices)
    stats = (_lib.SweepStats * max(1, n_sweeps))()
    spec = lda.component.spec(data.datum)
    check(lib().bias_lda_run_multi(ptr(dev), len(dev), C.c_uint64(seed), C.byref(spec), C.byref(data.c), data.n_groups,
                                   ptr(data.group_ptr), ptr(flat), lda.KK, lda.aa, n_sweeps, stats))
    _scatter_zz(flat, zz, True, data.group_ptr)
    return list(stats)[:n_sweeps]


def _store(filename: str, i: int, **arrays):
    """storesample (src/LDA.jl:45-64): the reference writes JLD (HDF5); no HDF5 tooling exists in
    this image, so samples are stored as .npz with the same keys."""
    name = f"{filename}{i}.npz" if filename.endswith("_") else f"{filename}_{i}.npz"
    np.savez(name, **arrays)


def collapsed_gibbs_sampler(model, xx, zz, n_burnins: int, n_lags: int, n_samples: int, *args,
                            store_every: int = 100, filename: str | None = None, ctx: Context | None = None,
                            verbose: bool = False, sample_hyperparam: bool = True, n_internals: int = 10,
                            KK_list: Sequence[int] = (), KK_dict: dict | None = None):
    """collapsed_gibbs_sampler!(model, xx, zz, n_burnins, n_lags, n_samples, ...).

    LDA / BMM: mutates zz in place and returns None (src/LDA.jl:67-148, src/BMM.jl:46-130).
    HDP: returns (KK_list, KK_dict, betas, gammas, alphas) (src/HDP.jl:398) and updates
    hdp.KK / hdp.gg / hdp.aa.  Positional extras follow the reference's order
    (HDP: sample_hyperparam, n_internals, store_every, filename; LDA/BMM: store_every, filename).
    filename=None disables sample files.
    """
    if isinstance(model, HDP):
        names = ["sample_hyperparam", "n_internals", "store_every", "filename"]
    else:
        names = ["store_every", "filename"]
    if len(args) > len(names):
        raise TypeError("too many positional arguments")
    loc = {"sample_hyperparam": sample_hyperparam, "n_internals": n_internals, "store_every": store_every,
           "filename": filename}
    loc.update(dict(zip(names, args)))
    sample_hyperparam, n_internals = loc["sample_hyperparam"], loc["n_internals"]
    store_every, filename = loc["store_every"], loc["filename"]

    ctx = ctx or default_context()
    grouped = not isinstance(model, BMM)
    data = FlatData(model.component, xx, grouped)
    n_iterations = n_burnins + n_samples * (n_lags + 1)                 # src/LDA.jl:82
    try:
        rs = ResidentSampler(ctx, model, data, _flat_zz(zz, grouped), sample_hyperparam=sample_hyperparam,
                             n_internals=n_internals)
    except _lib.BiasError as e:
        if e.code == _lib.EINVAL:
            raise ValueError(str(e)) from None                          # the reference's ArgumentError
        raise

    is_hdp = isinstance(model, HDP)
    if is_hdp:
        n_old = len(KK_list)
        KK_list = list(KK_list) + [0] * n_samples
        KK_dict = {} if (KK_dict is None or n_old == 0) else KK_dict
        betas, gammas, alphas = [None] * n_samples, np.zeros(n_samples), np.zeros(n_samples)
    try:
        for iteration in range(1, n_iterations + 1):
            rs.sweep(1)
            if verbose:
                s = rs.stats()
                burn = "Burning... " if iteration < n_burnins else ""
                print(f"{burn}iteration: {iteration}, KK={s.KK}, likelihood={s.log_likelihood:.2f}")
            if (iteration - n_burnins) % (n_lags + 1) == 0 and iteration > n_burnins:
                i = (iteration - n_burnins) // (n_lags + 1)
                if is_hdp:
                    hp, beta = rs.hdp_state()
                    flat = rs.get_zz()
                    _scatter_zz(flat, zz, True, data.group_ptr)
                    KK_list[n_old + i - 1] = hp.KK
                    KK_dict[hp.KK] = [np.array(g, dtype=np.int64) for g in zz]
                    alphas[i - 1], gammas[i - 1], betas[i - 1] = hp.aa, hp.gg, beta
                if filename is not None and i % store_every == 0:
                    _store(filename, i, zz=rs.get_zz(), iteration=iteration)
        flat = rs.get_zz()
        _scatter_zz(flat, zz, grouped, data.group_ptr)
        if is_hdp:
            hp, beta = rs.hdp_state()
            model.KK, model.gg, model.aa = hp.KK, hp.gg, hp.aa
            if n_samples > 0:                                           # src/HDP.jl:390-395
                KK_list[n_old + n_samples - 1] = hp.KK
                KK_dict[hp.KK] = [np.array(g, dtype=np.int64) for g in zz]
                alphas[n_samples - 1], gammas[n_samples - 1], betas[n_samples - 1] = hp.aa, hp.gg, beta
        if filename is not None:
            _store(filename, (n_iterations - n_burnins) // (n_lags + 1), zz=flat, iteration=n_iterations)
    finally:
        rs.close()
    if is_hdp:
        return KK_list, KK_dict, betas, gammas, alphas
    return None


def RCRP_gibbs_sampler(rcrp: RCRP, xx, zz, n_burnins: int, n_lags: int, n_samples: int,
                       sample_hyperparam: bool = True, n_internals: int = 10, store_every: int = 100,
                       filename: str | None = None, KK_list: Sequence[int] = (), KK_dict: dict | None = None,
                       ctx: Context | None = None, verbose: bool = False):
    """RCRP_gibbs_sampler!(rcrp, xx, zz, n_burnins, n_lags, n_samples, sample_hyperparam, n_internals,
    store_every, filename, KK_list, KK_dict) -> (KK_list, KK_dict)   (src/RCRP.jl:77-416).
    xx / zz are lists over epochs; zz, rcrp.KK and rcrp.aa are updated in place."""
    ctx = ctx or default_context()
    data = FlatData(rcrp.component, xx, True)
    n_iterations = n_burnins + n_samples * (n_lags + 1)                 # src/RCRP.jl:159
    n_old = len(KK_list)
    KK_list = list(KK_list) + [0] * n_samples
    KK_dict = {} if (KK_dict is None or n_old == 0) else KK_dict
    try:
        rs = ResidentSampler(ctx, rcrp, data, _flat_zz(zz, True), sample_hyperparam=sample_hyperparam,
                             n_internals=n_internals)
    except _lib.BiasError as e:
        if e.code == _lib.EINVAL:
            raise ValueError(str(e)) from None
        raise
    try:
        if n_samples > 0:
            KK_list[0] = rs.current_KK()                                # :147 (KK after the inactive-chain prune)
        for iteration in range(1, n_iterations + 1):
            if iteration > n_burnins:                                   # :166-168
                rs.set_sample_hyperparam(False)
            rs.sweep(1)
            if verbose:
                s = rs.stats()
                burn = "Burning... " if iteration < n_burnins else ""
                print(f"{burn}iteration: {iteration}, KK={s.KK}, likelihood={s.log_likelihood:.2f}")
            if (iteration - n_burnins) % (n_lags + 1) == 0 and iteration > n_burnins:
                i = n_old + (iteration - n_burnins) // (n_lags + 1)
                KK = rs.current_KK()
                _scatter_zz(rs.get_zz(), zz, True, data.group_ptr)
                KK_list[i - 1] = KK
                KK_dict[KK] = [np.array(g, dtype=np.int64) for g in zz]
                if filename is not None and i % store_every == 0:
                    _store(filename, i, zz=np.concatenate(zz), KK_list=np.asarray(KK_list), iteration=iteration)
        _scatter_zz(rs.get_zz(), zz, True, data.group_ptr)
        rp, aa = rs.rcrp_state()
        rcrp.KK, rcrp.aa = int(rp.KK), aa
    finally:
        rs.close()
    return KK_list, KK_dict


def CRF_gibbs_sampler(hdp: HDP, xx, zz, n_burnins: int, n_lags: int, n_samples: int, sample_hyperparam: bool = True,
                      n_internals: int = 10, store_every: int = 100, filename: str | None = None,
                      KK_list: Sequence[int] = (), KK_dict: dict | None = None, ctx: Context | None = None,
                      verbose: bool = False):
    """CRF_gibbs_sampler!(hdp, xx, zz, n_burnins, n_lags, n_samples, sample_hyperparam, n_internals, store_every,
    filename, KK_list, KK_dict) -> (KK_list, KK_dict)   (src/HDP.jl:402-709): the HDP sampled through the Chinese
    restaurant franchise (a table for every customer, a dish for every table).  zz, hdp.KK, hdp.gg and hdp.aa are
    updated in place."""
    ctx = ctx or default_context()
    data = FlatData(hdp.component, xx, True)
    n_iterations = n_burnins + n_samples * (n_lags + 1)
    n_old = len(KK_list)
    KK_list = list(KK_list) + [0] * n_samples
    KK_dict = {} if (KK_dict is None or n_old == 0) else KK_dict
    try:
        rs = ResidentSampler(ctx, hdp, data, _flat_zz(zz, True), sample_hyperparam=sample_hyperparam,
                             n_internals=n_internals, crf=True)
    except _lib.BiasError as e:
        if e.code == _lib.EINVAL:
            raise ValueError(str(e)) from None
        raise
    try:
        if n_samples > 0:
            KK_list[0] = hdp.KK                                         # :468
        for iteration in range(1, n_iterations + 1):
            rs.sweep(1)
            if verbose:
                s = rs.stats()
                burn = "Burning... " if iteration < n_burnins else ""
                print(f"{burn}iteration: {iteration}, KK={s.KK}, gg={s.gg:.2f}, aa={s.aa:.2f}, likelihood={s.log_likelihood:.2f}")
            if (iteration - n_burnins) % (n_lags + 1) == 0 and iteration > n_burnins:
                i = n_old + (iteration - n_burnins) // (n_lags + 1)
                hp, _ = rs.hdp_state()
                _scatter_zz(rs.get_zz(), zz, True, data.group_ptr)
                KK_list[i - 1] = hp.KK
                KK_dict[hp.KK] = [np.array(g, dtype=np.int64) for g in zz]
                if filename is not None and i % store_every == 0:
                    _store(filename, i, zz=np.concatenate(zz), KK_list=np.asarray(KK_list), iteration=iteration)
        _scatter_zz(rs.get_zz(), zz, True, data.group_ptr)
        hp, _ = rs.hdp_state()
        hdp.KK, hdp.gg, hdp.aa = hp.KK, hp.gg, hp.aa
        if n_samples > 0:                                               # :697-701
            KK_list[n_old + n_samples - 1] = hp.KK
            KK_dict[hp.KK] = [np.array(g, dtype=np.int64) for g in zz]
    finally:
        rs.close()
    return KK_list, KK_dict


def _flatten_epochs(xx3):
    """xx::Vector{Vector{Vector{T}}} (epochs of restaurants of customers) -> (list of restaurants, epoch_ptr)"""
    groups = [g for ep in xx3 for g in ep]
    eptr = np.concatenate([[0], np.cumsum([len(ep) for ep in xx3])]).astype(np.int64)
    return groups, eptr


def RCRF_gibbs_sampler(rcrf: RCRF, xx, zz, n_burnins: int, n_lags: int, n_samples: int, sample_hyperparam: bool = True,
                       n_internals: int = 10, store_every: int = 100, filename: str | None = None, join_tables: bool = True,
                       KK_list: Sequence[int] = (), KK_dict: dict | None = None, ctx: Context | None = None,
                       verbose: bool = False):
    """RCRF_gibbs_sampler!(rcrf, xx, zz, n_burnins, n_lags, n_samples, sample_hyperparam, n_internals, store_every,
    filename, join_tables, KK_list, KK_dict) -> (KK_list, KK_dict)   (src/RCRF.jl:122-938).  xx / zz are lists over
    epochs of lists over restaurants; zz, rcrf.KK, rcrf.gg and rcrf.aa are updated in place."""
    ctx = ctx or default_context()
    groups, eptr = _flatten_epochs(xx)
    zgroups, _ = _flatten_epochs(zz)
    data = FlatData(rcrf.component, groups, True)
    n_iterations = n_burnins + n_samples * (n_lags + 1)
    n_old = len(KK_list)
    KK_list = list(KK_list) + [0] * n_samples
    KK_dict = {} if (KK_dict is None or n_old == 0) else KK_dict
    try:
        rs = ResidentSampler(ctx, rcrf, data, _flat_zz(zgroups, True), sample_hyperparam=sample_hyperparam,
                             n_internals=n_internals, epoch_ptr=eptr, join_tables=join_tables)
    except _lib.BiasError as e:
        if e.code == _lib.EINVAL:
            raise ValueError(str(e)) from None
        raise

    def snapshot():
        _scatter_zz(rs.get_zz(), zgroups, True, data.group_ptr)
        it = iter(zgroups)
        for ep in zz:
            for j in range(len(ep)):
                ep[j] = next(it)
        return [[np.array(g, dtype=np.int64) for g in ep] for ep in zz]

    try:
        if n_samples > 0:
            KK_list[0] = rcrf.KK
        for iteration in range(1, n_iterations + 1):
            if iteration > n_burnins:                                   # src/RCRF.jl:226-228
                rs.set_sample_hyperparam(False)
            rs.sweep(1)
            if verbose:
                s = rs.stats()
                print(f"iteration: {iteration}, KK={s.KK}, likelihood={s.log_likelihood:.2f}")
            if (iteration - n_burnins) % (n_lags + 1) == 0 and iteration > n_burnins:
                i = n_old + (iteration - n_burnins) // (n_lags + 1)
                KK = rs.current_KK()
                KK_list[i - 1] = KK
                KK_dict[KK] = snapshot()
                if filename is not None and i % store_every == 0:
                    _store(filename, i, zz=rs.get_zz(), KK_list=np.asarray(KK_list), iteration=iteration)
        snapshot()
        rcrf.KK, rcrf.gg, rcrf.aa = rs.rcrf_state()
    finally:
        rs.close()
    return KK_list, KK_dict


def posterior(model, xx, zz, KK: int | None = None, ctx: Context | None = None):
    """posterior(model, xx, zz) -> ([posterior(component_k)], nn)  (src/LDA.jl:151-167, src/BMM.jl:134-151).
    RCRP: posterior(rcrp, xx, KK_dict, KK) (src/RCRP.jl:418-453).
    The sufficient statistics are accumulated on the device by the initialisation pass."""
    from .distributions import Dirichlet, Gaussian1D
    ctx = ctx or default_context()
    if isinstance(model, RCRP):
        if KK is None:
            raise TypeError("posterior(rcrp, xx, KK_dict, KK): KK is required")
        zz = zz[KK]
        sized = RCRP(model.component, KK, model.aa, None, model.a1, model.a2, max(model.Kmax, KK + 2))
        model = sized
    grouped = not isinstance(model, BMM)
    data = FlatData(model.component, xx, grouped)
    rs = ResidentSampler(ctx, model, data, _flat_zz(zz, grouped))
    try:
        comp = rs.components()
        c0 = model.component
        posts = []
        for k in range(rs.current_KK()):
            if isinstance(c0, MultinomialDirichlet):
                posts.append(Dirichlet(comp["mi"][k] + c0.aa))
            else:
                n = int(comp["nn"][k])
                if n > 0:
                    denom = c0.vv + n * c0.v0
                    posts.append(Gaussian1D((n * c0.v0 * comp["mu"][k] + c0.vv * c0.m0) / denom, c0.v0 * c0.vv / denom))
                else:
                    posts.append(Gaussian1D(c0.m0, c0.v0))
        nn = rs.group_counts() if grouped else comp["nn"]
    finally:
        rs.close()
    return posts, nn


# ------------------------------------------------------------------ HDP verification helpers
def hdp_hyper_aux(ctx: Context, group_ptr, aa: float, sweep: int = 0, internal: int = 0, uniforms=None):
    """Device part of sample_hyperparam! (src/HDP.jl:133-144) -> (sum_j log w_j, sum_j s_j)."""
    gp = _lib.as_i64(group_ptr)
    u = None if uniforms is None else _lib.as_f64(uniforms)
    out = np.zeros(2)
    check(lib().bias_hdp_hyper_aux(ctx.h, gp.size - 1, ptr(gp), aa, sweep, internal, ptr(u), ptr(out)))
    return float(out[0]), float(out[1])


def hdp_hyper_draws(seed, a1, a2, g1, g2, n_tables, sum_logw, sum_s, KK, aa, gg, n):
    """Host part of sample_hyperparam! (src/HDP.jl:146-158): n draws of (alpha0, gamma) from the same conditioning values."""
    a, g = np.zeros(n), np.zeros(n)
    check(lib().bias_hdp_hyper_draws(C.c_uint64(seed), a1, a2, g1, g2, n_tables, sum_logw, sum_s, KK, aa, gg, n, ptr(a), ptr(g)))
    return a, g


def hdp_beta_draws(seed, tables, gg, n):
    """n draws of my_beta ~ Dirichlet([sum_j M_jk ..., gg]) (src/HDP.jl:362-364) -> [n, KK + 1]."""
    t = _lib.as_i64(tables)
    out = np.zeros((n, t.size + 1))
    check(lib().bias_hdp_beta_draws(C.c_uint64(seed), t.size, ptr(t), gg, n, ptr(out)))
    return out


def stirling_rows(ctx: Context, n_rows: int) -> np.ndarray:
    """Rows 1..n_rows of the device-resident normalised log-Stirling table (packed lower triangle)."""
    out = np.zeros(n_rows * (n_rows + 1) // 2)
    check(lib().bias_stirling_rows(ctx.h, n_rows, ptr(out)))
    return out


def hdp_table_conditionals(ctx: Context, n, a, uniforms):
    """p(m | n, a) for each query and the draw for the supplied uniform (src/HDP.jl:353-358)."""
    n, a, u = _lib.as_i32(n), _lib.as_f64(a), _lib.as_f64(uniforms)
    n_max = int(n.max()) if len(n) else 1
    prob = np.zeros((len(n), n_max))
    draws = np.zeros(len(n), dtype=np.int32)
    check(lib().bias_hdp_table_conditionals(ctx.h, len(n), ptr(n), ptr(a), ptr(u), n_max, ptr(prob), ptr(draws)))
    return prob, draws
