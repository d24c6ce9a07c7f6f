"""ctypes loader for oracle/libko_oracle.so (C restatement, see ccd_oracle.c).
TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
METHODS = {"mvr": 0, "maxent": 1, "sdp_ccd": 2}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libko_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ccd_oracle.c",)]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libko_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        _LIB.ko_solve_ccd.restype = C.c_int
        _LIB.ko_solve_ccd.argtypes = [C.c_int, dp, C.c_int64, C.c_double, C.c_int, C.c_double, C.c_double,
                                      C.c_double, C.c_double, C.c_int, dp, dp, C.POINTER(C.c_int), dp, dp, C.c_int, C.c_int64, dp, C.c_int64]
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def solve_ccd(Sigma, method, m=1, niter=100, tol=1e-6, lambda_min=1e-6, sdp_lambda=0.5, sdp_mu=0.8,
              robust=False, s_init=None, max_sweeps=0, max_coords=0, lapack_init=False, coord_stride=1):
    """Returns dict(s, sweeps, trace, touched_entries, n_updates, rc).  coord_stride > 1: timing sample only (every
    coord_stride-th coordinate of a sweep; s is then not a solution), see ccd_oracle.c."""
    Sigma = np.asfortranarray(Sigma, dtype=np.float64)
    p = Sigma.shape[0]
    s = np.zeros(p)
    trace = np.zeros(max(niter, 1))
    stats = np.zeros(3)
    U0 = None
    if lapack_init:
        # the reference's `cholesky(Symmetric(...))` is LAPACK dpotrf (utilities.jl:140, :237, :309)
        import scipy.linalg as sla
        kap = (m + 1) / m
        A = kap * (np.triu(Sigma) + np.triu(Sigma, 1).T)
        if method != "sdp_ccd":
            A[np.diag_indices(p)] += lambda_min - np.asarray(s_init, dtype=np.float64)
        U0 = np.ascontiguousarray(sla.cholesky(A, lower=False, check_finite=False))
    sw = C.c_int(0)
    si = None if s_init is None else np.ascontiguousarray(s_init, dtype=np.float64)
    rc = lib().ko_solve_ccd(METHODS[method], _dp(Sigma), p, float(m), int(niter), float(tol), float(lambda_min),
                            float(sdp_lambda), float(sdp_mu), int(robust), _dp(si), _dp(s), C.byref(sw),
                            _dp(trace), _dp(stats), int(max_sweeps), int(max_coords), _dp(U0), int(coord_stride))
    return dict(s=s, sweeps=sw.value, trace=trace[:sw.value].copy(), touched_entries=stats[0],
                n_updates=stats[1], loop_seconds=stats[2], rc=rc)
