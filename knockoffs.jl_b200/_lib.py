"""ctypes binding of libknockoffs_b200.so -- exactly the symbols include/knockoffs_b200.h declares.

There is NO CPU fallback: if the shared library is missing, or no CUDA device is present,
every entry point raises.  (The CPU oracle lives in oracle/ and is test infrastructure only.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KB_LIB_PATH") or os.path.join(_HERE, "libknockoffs_b200.so")   # KB_LIB_PATH: A/B builds

KB_OK, KB_ERR_INVALID, KB_ERR_NOT_PD, KB_ERR_NAN, KB_ERR_CUDA = 0, -1, -2, -3, -4
KB_MAX_TRACE = 128
METHODS = {"mvr": 0, "maxent": 1, "sdp_ccd": 2, "equi": 3}

c_double_p = C.POINTER(C.c_double)


class SolveOpts(C.Structure):
    _fields_ = [("niter", C.c_int32), ("tol", C.c_double), ("lambda_min", C.c_double), ("sdp_lambda", C.c_double),
                ("sdp_mu", C.c_double), ("robust", C.c_int32), ("verbose", C.c_int32), ("mode", C.c_int32),
                ("refresh_every", C.c_int32), ("objective", C.c_int32)]


class SolveInfo(C.Structure):
    _fields_ = [("sweeps", C.c_int32), ("status", C.c_int32), ("mode", C.c_int32), ("launches", C.c_int32),
                ("trace", C.c_double * KB_MAX_TRACE), ("solve_ms", C.c_double), ("init_ms", C.c_double),
                ("ref_bytes", C.c_double), ("touched_bytes", C.c_double), ("flops", C.c_double),
                ("updates", C.c_double), ("objective", C.c_double)]


class GroupOpts(C.Structure):
    _fields_ = [("outer_iter", C.c_int32), ("inner_pca_iter", C.c_int32), ("inner_ccd_iter", C.c_int32),
                ("tol", C.c_double), ("eps", C.c_double), ("robust", C.c_int32), ("verbose", C.c_int32)]


class GroupInfo(C.Structure):
    _fields_ = [("outer_iters", C.c_int32), ("inner_sweeps", C.c_int32), ("n_directions", C.c_int32),
                ("status", C.c_int32), ("obj_init", C.c_double), ("obj", C.c_double), ("gamma_equi", C.c_double),
                ("trace_obj", C.c_double * KB_MAX_TRACE), ("trace_delta", C.c_double * KB_MAX_TRACE),
                ("trace_kind", C.c_int32 * KB_MAX_TRACE), ("steps", C.c_double), ("accepted", C.c_double),
                ("solve_ms", C.c_double), ("touched_bytes", C.c_double)]


# name -> (restype, argtypes); mirrors include/knockoffs_b200.h one to one
_vp, _i32, _i64, _u64, _d = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
SIGNATURES = {
    "kb_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "kb_create_background": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "kb_destroy": (None, [_vp]),
    "kb_last_error": (C.c_char_p, [_vp]),
    "kb_version": (C.c_char_p, []),
    "kb_default_solve_opts": (None, [C.POINTER(SolveOpts)]),
    "kb_launch_count": (_u64, [_vp]),
    "kb_stream": (_vp, [_vp]),
    "kb_aux_sms": (_i32, [_vp]),
    "kb_synchronize": (C.c_int, [_vp]),
    "kb_set_positive_cholesky": (C.c_int, [_vp, _i32]),
    "kb_positive_modified": (_i32, [_vp]),
    "kb_trim": (C.c_int, [_vp]),
    "kb_set_sample_chunk_rows": (C.c_int, [_vp, _i64]),
    "kb_solve_s": (C.c_int, [_vp, _vp, _i64, _i32, _d, C.POINTER(SolveOpts), _vp, _vp, C.POINTER(SolveInfo)]),
    "kb_solve_s_dev": (C.c_int, [_vp, _vp, _i64, _i32, _d, C.POINTER(SolveOpts), _vp, _vp, C.POINTER(SolveInfo)]),
    "kb_eigmin": (C.c_int, [_vp, _vp, _i64, c_double_p]),
    "kb_default_group_opts": (None, [C.POINTER(GroupOpts)]),
    "kb_solve_s_group": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _d, C.POINTER(GroupOpts), _vp, _i64, _vp, c_double_p,
                                   C.POINTER(GroupInfo)]),
    "kb_solve_s_group_sharded": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _d, C.POINTER(GroupOpts), _vp, _i64, _vp, _vp,
                                           _vp, c_double_p, C.POINTER(GroupInfo)]),
    "kb_modelx_gaussian_group_knockoffs": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _i32,
                                                     C.POINTER(GroupOpts), _vp, _i64, _vp, _u64, _i64, _vp, _vp,
                                                     c_double_p, C.POINTER(GroupInfo)]),
    "kb_gram": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _d, _vp, _vp]),
    "kb_gram_partial_dev": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "kb_gram_finish_dev": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i32, _d]),
    "kb_cov_shrink_lw": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, c_double_p]),
    "kb_cond_factor": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp]),
    "kb_cond_factor_dev": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp]),
    "kb_cond_factor_block_dev": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp]),
    "kb_sample": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _vp, _u64, _i64, _vp]),
    "kb_condition": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _u64, _i64, _vp]),
    "kb_sample_dev": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _vp, _i64, _u64, _i64, _vp, _i64]),
    "kb_modelx_gaussian_knockoffs": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i32, _i32, C.POINTER(SolveOpts), _vp,
                                               _vp, _u64, _i64, _vp, _vp, C.POINTER(SolveInfo)]),
    "kb_marginal_stats": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp]),
    "kb_marginal_stats_dev": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i32, _vp]),
    "kb_mk_statistics": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "kb_threshold": (C.c_int, [_vp, _vp, _i64, _d, _i32, _i64, c_double_p]),
    "kb_mk_threshold": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _d, _i64, c_double_p]),
    "kb_marginal_filter": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp,
                                     _vp, _vp, C.POINTER(_i64)]),
    "kb_fit_marginal": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _i32, C.POINTER(SolveOpts), _vp, _i32,
                                  _i32, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(SolveInfo)]),
    "kb_approx_modelx_gaussian_knockoffs": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i32, _i32, _i32, C.POINTER(SolveOpts),
                                                      _vp, _u64, _vp, _vp, _vp, _vp, c_double_p, C.POINTER(SolveInfo)]),
    "kb_approx_solve_windows": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i32, _i32, _i32, _i32, _i32, C.POINTER(SolveOpts),
                                          _vp, _vp, _vp, c_double_p, C.POINTER(SolveInfo)]),
    "kb_approx_condition_windows": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _u64,
                                              _i64, _vp]),
    "kb_feasible_gamma": (C.c_int, [_vp, _vp, _i64, _vp, _d, c_double_p]),
    "kb_choose_group_reps": (C.c_int, [_vp, _vp, _i64, _vp, _d, _vp, _i64, _vp, C.POINTER(_i64)]),
    "kb_cond_indep_corr": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "kb_solve_s_graphical_group": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _d, C.POINTER(GroupOpts), _vp, _i64, _vp,
                                             _vp, c_double_p, C.POINTER(GroupInfo)]),
    "kb_modelx_gaussian_rep_group_knockoffs": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _d, _i32,
                                                         C.POINTER(GroupOpts), _vp, _i64, _vp, _i64, _vp, _u64, _vp, _vp,
                                                         C.POINTER(_i64), _vp, _vp, c_double_p, C.POINTER(GroupInfo)]),
    "kb_create_multi": (C.c_int, [C.POINTER(_vp), C.POINTER(C.c_int), _i32]),
    "kb_destroy_multi": (None, [_vp]),
    "kb_multi_size": (_i32, [_vp]),
    "kb_multi_ctx": (_vp, [_vp, _i32]),
    "kb_multi_last_error": (C.c_char_p, [_vp]),
    "kb_multi_gram": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _d, _vp, _vp]),
    "kb_multi_condition": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _u64, _vp]),
    "kb_multi_modelx_gaussian_knockoffs": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i32, _i32, C.POINTER(SolveOpts), _vp, _vp,
                                                     _u64, _vp, _vp, C.POINTER(SolveInfo)]),
    "kb_multi_solve_s_many": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, C.POINTER(SolveOpts), _vp, _vp]),
    "kb_multi_solve_s_group_blockdiag": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _d, C.POINTER(GroupOpts), _vp, c_double_p, _vp]),
    "kb_test_dgemm": (C.c_int, [_vp, _i32, _i32, _i64, _i64, _i64, _d, _vp, _i64, _vp, _i64, _d, _vp, _i64]),
    "kb_test_dgemm_dev": (C.c_int, [_vp, _i32, _i32, _i64, _i64, _i64, _d, _vp, _i64, _vp, _i64, _d, _vp, _i64]),
    "kb_test_dgemm_tma_dev": (C.c_int, [_vp, _i64, _i64, _i64, _d, _vp, _i64, _vp, _i64, _d, _vp, _i64]),
    "kb_test_factor_band": (C.c_int, [_vp, _vp, _i64, _i32, C.POINTER(C.c_int32)]),
    "kb_test_potrf": (C.c_int, [_vp, _vp, _i64, _vp]),
    "kb_test_potrf_positive": (C.c_int, [_vp, _vp, _i64, _vp, _vp, C.POINTER(C.c_int32)]),
    "kb_test_spd_inverse": (C.c_int, [_vp, _vp, _i64, _vp]),
    "kb_test_randn": (C.c_int, [_vp, _u64, _i64, _i64, _i64, _vp]),
}

REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.c_int32)
KB_REDUCE_MIN, KB_REDUCE_SUM, KB_REDUCE_MAX = 0, 1, 2

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree shared library and set every prototype.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is not built (run `make` or `python -c 'import __graft_entry__ as g; g.build()'`); "
                "knockoffs_b200 has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)     # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class KnockoffsError(RuntimeError):
    """error(...) of the reference (invalid arguments, NaN/Inf delta) or a CUDA failure."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class PosDefException(KnockoffsError):
    """LinearAlgebra.PosDefException of the reference (initial cholesky / downdate)."""


def check(ctx_handle, rc: int):
    if rc == KB_OK:
        return
    msg = load().kb_last_error(ctx_handle).decode("utf-8", "replace")
    if rc == KB_ERR_NOT_PD:
        raise PosDefException(rc, msg)
    raise KnockoffsError(rc, msg)
