"""ctypes binding of libbias_cuda (include/bias_cuda.h).

This is the reference-side binding a Python caller uses; the Julia shim in
julia/BIASCuda.jl binds the same symbols with `ccall`.  There is no CPU
fallback: if the shared library is missing, loading fails loudly, and on a
machine without a B200 every compute entry point returns BIAS_ENODEV.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BIAS_CUDA_LIB: another build of the same library (kernel experiments); the default is the in-tree build
LIB_PATH = os.environ.get("BIAS_CUDA_LIB") or os.path.join(_HERE, "csrc", "libbias_cuda.so")

OK, EINVAL, ENODEV, ECUDA, ENOMEM, ERANGE, ESTATE = range(7)
MD, G1G1 = 0, 1
INT, SENT, F64 = 0, 1, 2
MODEL_LDA, MODEL_BMM, MODEL_HDP, MODEL_RCRP, MODEL_CRF, MODEL_RCRF = 0, 1, 2, 3, 4, 5

# every symbol include/bias_cuda.h declares (tests check the library exports all of them)
SYMBOLS = [
    "bias_version", "bias_last_error", "bias_device_count", "bias_ctx_create", "bias_ctx_destroy",
    "bias_ctx_sync", "bias_ctx_num_sms", "bias_ctx_launch_count", "bias_lda_run", "bias_bmm_run", "bias_hdp_run",
    "bias_model_create", "bias_model_destroy", "bias_model_sweep", "bias_model_set_zz", "bias_model_get_zz",
    "bias_model_last_stats", "bias_model_get_components", "bias_model_get_group_counts", "bias_model_get_hdp",
    "bias_model_conditionals", "bias_model_frozen_draws", "bias_hdp_table_conditionals", "bias_stirling_rows",
    "bias_model_log_likelihood", "bias_model_delta_begin", "bias_model_delta_begin_zero", "bias_model_delta_pack", "bias_model_delta_apply", "bias_model_last_sweep_ms",
    "bias_rcrp_run", "bias_rcrp_create", "bias_model_get_rcrp", "bias_model_set_sample_hyperparam", "bias_model_reload", "bias_model_reload_u16", "bias_model_get_zz_u16", "bias_model_heldout_log_likelihood",
    "bias_model_set_host_seed", "bias_hdp_shard_pass", "bias_hdp_shard_births", "bias_hdp_shard_adopt", "bias_hdp_shard_prune",
    "bias_hdp_shard_tables", "bias_hdp_shard_draw", "bias_hdp_shard_end",
    "bias_crf_sweep_with_uniforms", "bias_crf_get_tables", "bias_model_delta_pack16", "bias_model_delta_apply16",
    "bias_rcrf_create", "bias_model_get_rcrf",
    "bias_model_frozen_draws_dev", "bias_model_sweep_with_uniforms", "bias_lda_run_multi", "bias_lda_run_multi_timing", "bias_group_create", "bias_group_sweep",
    "bias_group_destroy", "bias_model_share_prepare", "bias_model_share_export", "bias_model_share_attach", "bias_model_replica_digest", "bias_model_last_exchange", "bias_model_narrow_info", "bias_hdp_hyper_aux", "bias_hdp_hyper_draws", "bias_hdp_beta_draws",
]


class BiasError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libbias_cuda error {code}: {msg}")
        self.code = code


class Conjugate(C.Structure):
    _fields_ = [("kind", C.c_int32), ("datum", C.c_int32), ("dd", C.c_int64), ("aa", C.c_double),
                ("m0", C.c_double), ("v0", C.c_double), ("vv", C.c_double)]


class Data(C.Structure):
    _fields_ = [("n_items", C.c_int64), ("word", C.c_void_p), ("x", C.c_void_p),
                ("sent_ptr", C.c_void_p), ("sent_word", C.c_void_p), ("sent_count", C.c_void_p)]


class SweepStats(C.Structure):
    _fields_ = [("log_likelihood", C.c_double), ("n_moved", C.c_int64), ("KK", C.c_int32),
                ("reserved", C.c_int32), ("gg", C.c_double), ("aa", C.c_double)]


class HdpParams(C.Structure):
    _fields_ = [("KK", C.c_int32), ("Kmax", C.c_int32), ("gg", C.c_double), ("g1", C.c_double),
                ("g2", C.c_double), ("aa", C.c_double), ("a1", C.c_double), ("a2", C.c_double),
                ("sample_hyperparam", C.c_int32), ("n_internals", C.c_int32)]


class RcrpParams(C.Structure):
    _fields_ = [("KK", C.c_int32), ("Kmax", C.c_int32), ("a1", C.c_double), ("a2", C.c_double),
                ("sample_hyperparam", C.c_int32), ("n_internals", C.c_int32)]


class RcrfParams(C.Structure):
    _fields_ = [("KK", C.c_int32), ("Kmax", C.c_int32), ("g1", C.c_double), ("g2", C.c_double), ("a1", C.c_double),
                ("a2", C.c_double), ("sample_hyperparam", C.c_int32), ("n_internals", C.c_int32), ("join_tables", C.c_int32)]


class ShareHandle(C.Structure):
    _fields_ = [("ipc", C.c_ubyte * 64), ("bytes", C.c_int64), ("dd", C.c_int64), ("Kpad", C.c_int32),
                ("hot_cap", C.c_int32), ("rank", C.c_int32), ("device", C.c_int32)]


_lib = None


def lib():
    """Load libbias_cuda.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). bias.jl_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pp = C.POINTER(C.c_void_p)
    L.bias_version.restype = C.c_int
    L.bias_last_error.restype = C.c_char_p
    L.bias_device_count.argtypes = [C.POINTER(C.c_int)]
    L.bias_ctx_create.argtypes = [C.c_int, C.c_uint64, vp, pp]
    L.bias_ctx_destroy.argtypes = [vp]
    L.bias_ctx_sync.argtypes = [vp]
    L.bias_ctx_num_sms.argtypes = [vp, C.POINTER(C.c_int)]
    L.bias_ctx_launch_count.argtypes = [vp, C.POINTER(i64)]
    cj, dt, st, hp = C.POINTER(Conjugate), C.POINTER(Data), C.POINTER(SweepStats), C.POINTER(HdpParams)
    L.bias_lda_run.argtypes = [vp, cj, dt, i64, vp, vp, i32, f64, i32, st]
    L.bias_bmm_run.argtypes = [vp, cj, dt, vp, i32, f64, i32, st]
    L.bias_hdp_run.argtypes = [vp, cj, dt, i64, vp, vp, hp, i32, vp, st]
    L.bias_model_create.argtypes = [vp, i32, cj, dt, i64, vp, vp, i32, f64, hp, pp]
    L.bias_model_destroy.argtypes = [vp]
    L.bias_model_sweep.argtypes = [vp, i32]
    L.bias_model_set_zz.argtypes = [vp, vp]
    L.bias_model_get_zz.argtypes = [vp, vp]
    L.bias_model_last_stats.argtypes = [vp, st]
    L.bias_model_get_components.argtypes = [vp, vp, vp, vp, vp]
    L.bias_model_get_group_counts.argtypes = [vp, vp]
    L.bias_model_get_hdp.argtypes = [vp, hp, vp]
    L.bias_model_conditionals.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.bias_model_frozen_draws.argtypes = [vp, vp, vp]
    L.bias_hdp_table_conditionals.argtypes = [vp, i64, vp, vp, vp, i32, vp, vp]
    L.bias_stirling_rows.argtypes = [vp, i32, vp]
    L.bias_model_log_likelihood.argtypes = [vp, C.POINTER(f64), C.POINTER(i64)]
    L.bias_model_delta_begin.argtypes = [vp]
    L.bias_model_delta_begin_zero.argtypes = [vp]
    L.bias_model_delta_pack.argtypes = [vp, pp, C.POINTER(i64), pp, C.POINTER(i64)]
    L.bias_model_delta_apply.argtypes = [vp]
    L.bias_model_last_sweep_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
    rp = C.POINTER(RcrpParams)
    L.bias_rcrp_run.argtypes = [vp, cj, dt, i64, vp, vp, rp, vp, i32, i32, st]
    L.bias_rcrp_create.argtypes = [vp, cj, dt, i64, vp, vp, rp, vp, pp]
    L.bias_model_get_rcrp.argtypes = [vp, rp, vp]
    L.bias_model_set_sample_hyperparam.argtypes = [vp, i32]
    L.bias_model_reload.argtypes = [vp, dt, i64, vp, vp]
    L.bias_model_reload_u16.argtypes = [vp, i64, vp, i64, vp, vp]
    L.bias_model_get_zz_u16.argtypes = [vp, vp]
    L.bias_model_heldout_log_likelihood.argtypes = [vp, vp, vp, C.POINTER(f64), C.POINTER(i64)]
    L.bias_model_set_host_seed.argtypes = [vp, C.c_uint64]
    L.bias_hdp_shard_pass.argtypes = [vp]
    L.bias_hdp_shard_births.argtypes = [vp, C.POINTER(i32), vp]
    L.bias_hdp_shard_adopt.argtypes = [vp, i32, vp]
    L.bias_hdp_shard_prune.argtypes = [vp]
    L.bias_hdp_shard_tables.argtypes = [vp, i32, pp, C.POINTER(i64), pp]
    L.bias_hdp_shard_draw.argtypes = [vp]
    L.bias_hdp_shard_end.argtypes = [vp]
    L.bias_crf_sweep_with_uniforms.argtypes = [vp, vp]
    L.bias_crf_get_tables.argtypes = [vp, vp, vp, vp, vp, vp]
    L.bias_model_delta_pack16.argtypes = [vp, i32, pp, C.POINTER(i64), pp, C.POINTER(i64)]
    L.bias_model_delta_apply16.argtypes = [vp]
    L.bias_rcrf_create.argtypes = [vp, cj, dt, i64, vp, i64, vp, vp, C.POINTER(RcrfParams), vp, vp, pp]
    L.bias_model_get_rcrf.argtypes = [vp, C.POINTER(i32), vp, vp]
    L.bias_model_frozen_draws_dev.argtypes = [vp, vp, vp]
    L.bias_model_sweep_with_uniforms.argtypes = [vp, vp]
    L.bias_lda_run_multi.argtypes = [vp, i32, C.c_uint64, cj, dt, i64, vp, vp, i32, f64, i32, st]
    L.bias_group_create.argtypes = [vp, i32, pp]
    L.bias_group_sweep.argtypes = [vp, i32]
    L.bias_group_destroy.argtypes = [vp]
    L.bias_model_share_prepare.argtypes = [vp, i32, i32, i64]
    L.bias_model_share_export.argtypes = [vp, C.POINTER(ShareHandle)]
    L.bias_model_share_attach.argtypes = [vp, C.POINTER(ShareHandle), i32]
    L.bias_model_replica_digest.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.bias_lda_run_multi_timing.argtypes = [vp]
    L.bias_model_last_exchange.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i64)]
    L.bias_hdp_hyper_aux.argtypes = [vp, i64, vp, f64, C.c_uint32, C.c_uint32, vp, vp]
    L.bias_hdp_hyper_draws.argtypes = [C.c_uint64, f64, f64, f64, f64, f64, f64, f64, i32, f64, f64, i32, vp, vp]
    L.bias_hdp_beta_draws.argtypes = [C.c_uint64, i32, vp, f64, i32, vp]
    L.bias_model_narrow_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    for name in SYMBOLS:
        if name != "bias_last_error":
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc: int):
    if rc != OK:
        raise BiasError(rc, lib().bias_last_error().decode("utf-8", "replace"))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().bias_device_count(C.byref(n))
    return n.value if rc == OK else 0


class Context:
    """bias_ctx: one CUDA device + stream + Philox seed."""

    def __init__(self, device: int = 0, seed: int = 0, stream: int | None = None):
        self.h = C.c_void_p()
        check(lib().bias_ctx_create(device, C.c_uint64(seed & (2**64 - 1)), C.c_void_p(stream or 0), C.byref(self.h)))
        self.device = device
        self.stream_handle = int(stream or 0)      # 0: the library created a private stream (stream 0 cannot be passed)

    def sync(self):
        check(lib().bias_ctx_sync(self.h))

    @property
    def num_sms(self) -> int:
        n = C.c_int()
        check(lib().bias_ctx_num_sms(self.h, C.byref(n)))
        return n.value

    @property
    def launches(self) -> int:
        n = C.c_int64()
        check(lib().bias_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def close(self):
        if self.h:
            lib().bias_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
