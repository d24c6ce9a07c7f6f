"""Host-side mirror of the reference's interface for the model-X Gaussian knockoff path.

Same names, argument meaning and error behaviour as Knockoffs.jl v2.0.2 (paths relative to that repo):
``solve_s`` (src/utilities.jl:27-51), ``modelX_gaussian_knockoffs`` (src/modelX.jl:39-66),
``condition`` (src/modelX.jl:96-122).  Every function is a thin ctypes call into
libknockoffs_b200.so; numpy arrays are passed as Fortran-ordered float64 (Julia ``Matrix{Float64}``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib
from ._lib import METHODS, GroupInfo, GroupOpts, SolveInfo, SolveOpts, check

SINGLE_KNOCKOFFS = ["mvr", "maxent", "equi", "sdp", "sdp_ccd"]   # src/Knockoffs.jl:99


def _f(a, name="array"):
    a = np.asarray(a)
    if a.dtype != np.float64 or not a.flags.f_contiguous:
        a = np.asfortranarray(a, dtype=np.float64)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _need(cond, what):
    """The reference raises DimensionMismatch / ArgumentError before touching data; here every array shape is checked
    before a raw pointer crosses the C ABI (the library trusts the sizes it is given)."""
    if not cond:
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, what)


def _square(A, name="Sigma"):
    _need(A.ndim == 2 and A.shape[0] == A.shape[1], f"{name} must be square but was {A.shape}")
    return A.shape[0]


class Context:
    """One per host thread / device (include/knockoffs_b200.h: kb_create)."""

    def __init__(self, device: int = 0, background: bool = False):
        """``background=True``: kb_create_background -- every kernel of this context is confined to the SM partition that
        leaves a few SMs free, for throughput work (sampling) that runs beside another context's solve on the same GPU."""
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = (self._lib.kb_create_background if background else self._lib.kb_create)(C.byref(h), int(device))
        if rc != 0:
            raise _lib.KnockoffsError(rc, f"kb_create(device={device}) failed: no usable CUDA device "
                                          "(knockoffs_b200 has no CPU fallback)")
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self._lib.kb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self._lib.kb_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self._lib.kb_stream(self.h) or 0)

    @property
    def aux_sms(self) -> int:
        """SMs the auxiliary stream is confined to (green context), 0 = no partition."""
        return int(self._lib.kb_aux_sms(self.h))

    def set_positive_cholesky(self, on: bool = True):
        """Opt in to the restated ``cholesky(Positive, ·)`` fallback of ``condition`` (parity unpinned, see the header)."""
        check(self.h, self._lib.kb_set_positive_cholesky(self.h, int(bool(on))))

    @property
    def positive_modified(self) -> int:
        return int(self._lib.kb_positive_modified(self.h))

    def set_sample_chunk_rows(self, rows: int):
        check(self.h, self._lib.kb_set_sample_chunk_rows(self.h, int(rows)))


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _method_id(method) -> int:
    name = str(method).lstrip(":")
    if name == "sdp":
        # src/utilities.jl:37-38 sends non-group :sdp to the Hypatia interior-point solver, which is out of
        # scope; the coordinate-descent SDP is :sdp_ccd.
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID,
                                  "method :sdp is the interior-point solver (not on the GPU path); use :sdp_ccd")
    if name not in METHODS:
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"Method must be one of {SINGLE_KNOCKOFFS} but was {name}")
    return METHODS[name]


def make_opts(niter=100, tol=1e-6, λmin=None, lambda_min=1e-6, λ=None, sdp_lambda=0.5, μ=None, sdp_mu=0.8,
              robust=False, verbose=False, refresh_every=-1, mode="auto", objective=True, **unknown) -> SolveOpts:
    """Keyword arguments of solve_MVR / solve_max_entropy / solve_sdp_ccd (utilities.jl:124-133, 221-230, 291-300)."""
    if unknown:
        raise TypeError(f"unsupported keyword argument(s): {sorted(unknown)}")   # Julia: MethodError
    o = SolveOpts()
    _lib.load().kb_default_solve_opts(C.byref(o))
    o.niter = int(niter)
    o.tol = float(tol)
    o.lambda_min = float(lambda_min if λmin is None else λmin)
    o.sdp_lambda = float(sdp_lambda if λ is None else λ)
    o.sdp_mu = float(sdp_mu if μ is None else μ)
    o.robust = int(bool(robust))
    o.verbose = int(bool(verbose))
    o.refresh_every = int(refresh_every)
    o.objective = int(bool(objective))
    o.mode = {"auto": 0, "stream": 1, "block": 2}[mode] if isinstance(mode, str) else int(mode)
    return o


def _info_dict(info: SolveInfo, method: str) -> dict:
    n = max(0, min(info.sweeps, _lib.KB_MAX_TRACE))
    return dict(sweeps=info.sweeps, status=info.status, launches=info.launches, trace=list(info.trace[:n]),
                solve_ms=info.solve_ms, init_ms=info.init_ms, ref_bytes=info.ref_bytes,
                touched_bytes=info.touched_bytes, updates=info.updates, flops=info.flops, objective=info.objective,
                method=method,
                mode={1: "stream", 2: "block"}.get(info.mode, "none"))


def _print_trace(info: dict):
    # what `verbose=true` prints in the reference (utilities.jl:167-170, 273-276, 313-316)
    for l, v in enumerate(info["trace"], 1):
        if info["method"] == "sdp_ccd":
            print(f"Iter {l}: sum(s) = {v}")
        else:
            print(f"Iter {l}: δ = {v}")


def solve_s(Sigma, method, m=1, s_init=None, ctx: Optional[Context] = None, return_info=False, **kwargs):
    """``solve_s(Σ::Symmetric, method; m, kwargs...)`` -- src/utilities.jl:27-51.  Only the upper
    triangle of ``Sigma`` is read.  Returns ``s`` (and the solver info dict if ``return_info``)."""
    ctx = ctx or default_context()
    Sigma = _f(Sigma)
    if Sigma.ndim != 2 or Sigma.shape[0] != Sigma.shape[1]:
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "Sigma must be square")
    p = Sigma.shape[0]
    mid = _method_id(method)
    opts = make_opts(**kwargs)
    info = SolveInfo()
    s = np.empty(p)
    si = None if s_init is None else np.ascontiguousarray(s_init, dtype=np.float64)
    _need(si is None or si.shape == (p,), f"s_init must have length {p}")
    _need(float(m) >= 1, f"m should be 1 or larger but was {m}.")
    rc = ctx._lib.kb_solve_s(ctx.h, _ptr(Sigma), p, mid, float(m), C.byref(opts), _ptr(si), _ptr(s), C.byref(info))
    check(ctx.h, rc)
    d = _info_dict(info, str(method).lstrip(":"))
    if opts.verbose:
        _print_trace(d)
    return (s, d) if return_info else s


def eigmin(Sigma, ctx: Optional[Context] = None) -> float:
    ctx = ctx or default_context()
    Sigma = _f(Sigma)
    p = _square(Sigma)
    out = C.c_double()
    check(ctx.h, ctx._lib.kb_eigmin(ctx.h, _ptr(Sigma), p, C.byref(out)))
    return out.value


def solve_equi(Sigma, m=1, ctx: Optional[Context] = None):
    """src/utilities.jl:104-111 (on a correlation matrix)."""
    return solve_s(Sigma, "equi", m=m, ctx=ctx)


def gram(X, center=True, denom=None, ctx: Optional[Context] = None):
    """(Xc'Xc)/denom and the column means -- the Gram inside ``cov(estimator, X)`` and
    ``mean(X, dims=1)`` of the 2-argument entry (src/modelX.jl:47-49)."""
    ctx = ctx or default_context()
    X = _f(X)
    _need(X.ndim == 2, "X must be a matrix")
    n, p = X.shape
    G = np.empty((p, p), order="F")
    mu = np.empty(p)
    check(ctx.h, ctx._lib.kb_gram(ctx.h, _ptr(X), n, p, int(bool(center)), float(denom or 0.0), _ptr(G), _ptr(mu)))
    return G, mu


def cov_shrink_lw(X, ctx: Optional[Context] = None):
    """``cov(LinearShrinkage(DiagonalUnequalVariance(), :lw), X)`` and ``vec(mean(X, dims=1))`` -- the default
    covariance_approximator of the 2-argument entry (src/modelX.jl:43-50).  Returns (Sigma, mu, lambda)."""
    ctx = ctx or default_context()
    X = _f(X)
    _need(X.ndim == 2, "X must be a matrix")
    n, p = X.shape
    Sig = np.empty((p, p), order="F")
    mu = np.empty(p)
    lam = C.c_double()
    check(ctx.h, ctx._lib.kb_cov_shrink_lw(ctx.h, _ptr(X), n, p, _ptr(Sig), _ptr(mu), C.byref(lam)))
    return Sig, mu, lam.value


def cond_factor(Sigma, S, m=1, ctx: Optional[Context] = None):
    """ΣinvS and the factor blocks of src/modelX.jl:106-120.  ``S`` is a vector (``Diagonal(s)``) or a
    dense p x p matrix.  Returns (SigmaInvS, Lblocks) with Lblocks of shape (p, p, 2m-1):
    [L_11, M_1, L_22, M_2, ..., L_mm] (block (i,k) of the pm x pm lower factor is L_kk if i == k, M_k if i > k)."""
    ctx = ctx or default_context()
    Sigma = _f(Sigma)
    p = _square(Sigma)
    S = np.asarray(S, dtype=np.float64)
    is_diag = S.ndim == 1
    S = np.ascontiguousarray(S) if is_diag else _f(S)
    _need(S.shape == ((p,) if is_diag else (p, p)), f"S must have length {p} or be {p} x {p}")
    m = int(m)
    if m < 1:
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"m should be 1 or larger but was {m}.")
    SiS = np.empty((p, p), order="F")
    Lb = np.empty((p, p, 2 * m - 1), order="F")
    check(ctx.h, ctx._lib.kb_cond_factor(ctx.h, _ptr(Sigma), p, _ptr(S), int(is_diag), m, _ptr(SiS), _ptr(Lb)))
    return SiS, Lb


def dense_factor_from_blocks(Lb: np.ndarray) -> np.ndarray:
    """The pm x pm lower factor of modelX.jl:120 assembled from the block form (tests / inspection)."""
    p = Lb.shape[0]
    m = (Lb.shape[2] + 1) // 2
    L = np.zeros((p * m, p * m))
    for k in range(m):
        L[k * p:(k + 1) * p, k * p:(k + 1) * p] = Lb[:, :, 2 * k]
        for i in range(k + 1, m):
            L[i * p:(i + 1) * p, k * p:(k + 1) * p] = Lb[:, :, 2 * k + 1]
    return L


def sample(X, mu, SigmaInvS, Lblocks, m=1, Z=None, seed=0, row_offset=0, ctx: Optional[Context] = None):
    """X̃ = repeat(X − (X .− μ')ΣinvS, 1, m) + Z·L (src/modelX.jl:112, 118-121).  ``Z`` (n x mp) is the
    injected ``randn(n, mp)``; ``Z=None`` draws it on the device (Philox4x32-10, keyed by seed/row/column)."""
    ctx = ctx or default_context()
    X = _f(X)
    _need(X.ndim == 2, "X must be a matrix")
    n, p = X.shape
    m = int(m)
    _need(m >= 1, f"m should be 1 or larger but was {m}.")
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    SiS = _f(SigmaInvS)
    Lb = _f(Lblocks)
    if Lb.ndim == 2 and m == 1:
        Lb = Lb.reshape(Lb.shape[0], Lb.shape[1], 1, order="F")
    _need(mu.shape == (p,), f"mu must have length {p}")
    _need(SiS.shape == (p, p), f"SigmaInvS must be {p} x {p}")
    _need(Lb.shape == (p, p, 2 * m - 1), f"Lblocks must be {p} x {p} x {2 * m - 1} ([L_11, M_1, ..., L_mm])")
    Zf = None if Z is None else _f(Z)
    if Zf is not None and Zf.shape != (n, m * p):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"Z must be {n} x {m * p}")
    Xko = np.empty((n, m * p), order="F")
    check(ctx.h, ctx._lib.kb_sample(ctx.h, _ptr(X), n, p, _ptr(mu), _ptr(SiS), _ptr(Lb), int(m), _ptr(Zf),
                                    int(seed), int(row_offset), _ptr(Xko)))
    return Xko


def condition(X, mu, Sigma, S, m=1, Z=None, seed=0, row_offset=0, ctx: Optional[Context] = None):
    """``condition(X, μ, Σ, S; m)`` -- src/modelX.jl:96-122, one library call (kb_condition): Σ⁻¹S and the factor
    blocks never leave the device.  ``S``: vector (``Diagonal(s)``) or dense p x p.  ``X`` may be a row shard
    (``row_offset`` = global index of its first row)."""
    ctx = ctx or default_context()
    m = int(m)
    if m < 1:
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"m should be 1 or larger but was {m}.")
    X = _f(X)
    n, p = X.shape
    Sigma = _f(Sigma)
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    is_diag = S.ndim == 1
    S = np.ascontiguousarray(S) if is_diag else _f(S)
    if Sigma.shape != (p, p) or mu.shape != (p,) or S.shape != ((p,) if is_diag else (p, p)):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "dimension mismatch between X, mu, Sigma and S")
    Zf = None if Z is None else _f(Z)
    if Zf is not None and Zf.shape != (n, m * p):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"Z must be {n} x {m * p}")
    Xko = np.empty((n, m * p), order="F")
    check(ctx.h, ctx._lib.kb_condition(ctx.h, _ptr(X), n, p, _ptr(mu), _ptr(Sigma), _ptr(S), int(is_diag), m, _ptr(Zf),
                                       int(seed), int(row_offset), _ptr(Xko)))
    return Xko


@dataclass
class GaussianKnockoff:
    """src/struct.jl:6-13 (field names as in the reference)."""
    X: np.ndarray
    Xko: np.ndarray
    s: np.ndarray
    Sigma: np.ndarray
    method: str
    m: int
    info: dict = field(default_factory=dict)


def modelX_gaussian_knockoffs(X, method, mu=None, Sigma=None, m=1, Z=None, seed=0, s_init=None,
                              ctx: Optional[Context] = None, **kwargs) -> GaussianKnockoff:
    """``modelX_gaussian_knockoffs(X, method, μ, Σ; m, kwargs...)`` -- src/modelX.jl:53-66: one library call
    (solve_s -> conditional factor -> sampling; intermediates never leave the device).
    With ``mu``/``Sigma`` omitted this is the 2-argument entry (src/modelX.jl:39-51): Σ̂ is the Ledoit-Wolf
    linear-shrinkage estimate (two device Grams) and μ̂ the column means."""
    ctx = ctx or default_context()
    X = _f(X)
    n, p = X.shape
    m = int(m)
    if m < 1:
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"m should be 1 or larger but was {m}.")
    if (mu is None) != (Sigma is None):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "give both mu and Sigma, or neither")
    if Sigma is None:
        Sigma, mu, _ = cov_shrink_lw(X, ctx=ctx)                    # modelX.jl:47-49
    Sigma = _f(Sigma)
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    if Sigma.shape != (p, p) or mu.shape != (p,):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "dimension mismatch between X, mu and Sigma")
    mid = _method_id(method)
    opts = make_opts(**kwargs)
    info = SolveInfo()
    Zf = None if Z is None else _f(Z)
    if Zf is not None and Zf.shape != (n, m * p):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"Z must be {n} x {m * p}")
    si = None if s_init is None else np.ascontiguousarray(s_init, dtype=np.float64)
    _need(si is None or si.shape == (p,), f"s_init must have length {p}")
    Xko = np.empty((n, m * p), order="F")
    s = np.empty(p)
    rc = ctx._lib.kb_modelx_gaussian_knockoffs(ctx.h, _ptr(X), n, p, _ptr(mu), _ptr(Sigma), mid, m, C.byref(opts),
                                               _ptr(si), _ptr(Zf), int(seed), 0, _ptr(Xko), _ptr(s), C.byref(info))
    check(ctx.h, rc)
    d = _info_dict(info, str(method).lstrip(":"))
    if opts.verbose:
        _print_trace(d)
    Ssym = np.triu(Sigma) + np.triu(Sigma, 1).T
    return GaussianKnockoff(X, Xko, s, Ssym, str(method).lstrip(":"), m, d)


# ---- group knockoffs ---------------------------------------------------------------------------------
GROUP_METHODS = {"mvr": 0, "maxent": 1, "sdp": 2, "equi": 3}      # group :sdp is coordinate descent (group.jl:403-404)


def make_group_opts(outer_iter=100, inner_pca_iter=1, inner_ccd_iter=1, tol=1e-4, ϵ=None, eps=1e-6, robust=False,
                    verbose=False, **unknown) -> GroupOpts:
    """Keyword arguments of solve_group_{max_entropy,mvr,sdp}_hybrid (group.jl:870-881, 951-962, 1057-1068)."""
    if unknown:
        raise TypeError(f"unsupported keyword argument(s): {sorted(unknown)}")
    o = GroupOpts()
    _lib.load().kb_default_group_opts(C.byref(o))
    o.outer_iter, o.inner_pca_iter, o.inner_ccd_iter = int(outer_iter), int(inner_pca_iter), int(inner_ccd_iter)
    o.tol = float(tol)
    o.eps = float(eps if ϵ is None else ϵ)
    o.robust, o.verbose = int(bool(robust)), int(bool(verbose))
    return o


def _group_info_dict(info: GroupInfo, method: str) -> dict:
    n = max(0, min(info.inner_sweeps, _lib.KB_MAX_TRACE))
    return dict(outer_iters=info.outer_iters, inner_sweeps=info.inner_sweeps, n_directions=info.n_directions,
                obj_init=info.obj_init, obj=info.obj, gamma_equi=info.gamma_equi,
                trace=[("pca" if info.trace_kind[i] == 0 else "ccd", info.trace_obj[i], info.trace_delta[i])
                       for i in range(n)],
                steps=info.steps, accepted=info.accepted, solve_ms=info.solve_ms, touched_bytes=info.touched_bytes,
                method=method)


def _group_method_id(method) -> int:
    name = str(method).lstrip(":")
    if name not in GROUP_METHODS:
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID,
                                  f"Method must be one of {sorted(GROUP_METHODS)} (the conic-solver group methods are "
                                  f"not on the GPU path) but was {name}")
    return GROUP_METHODS[name]


def _print_group_trace(d: dict):
    print(f"{d['method']} initial obj = {d['obj_init']}")
    for it, (kind, obj, delta) in enumerate(d["trace"], 1):
        print(f"Iter {it} ({kind.upper()}): obj = {obj}, δ = {delta}")


def solve_s_group(Sigma, groups, method, m=1, V=None, ctx: Optional[Context] = None, return_info=False, **kwargs):
    """``solve_s_group(Σ::Symmetric, groups, method; m, kwargs...)`` -- src/group.jl:348-422.
    Returns ``(S, gammas, obj)`` like the reference (``gammas`` is empty for the CCD methods).
    ``V`` (p x nv, caller's variable order) replaces get_PCA_vectors' direction set."""
    ctx = ctx or default_context()
    Sigma = _f(Sigma)
    p = _square(Sigma)
    groups = np.ascontiguousarray(groups, dtype=np.int64)
    if groups.shape != (p,):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "Length of groups should be equal to dimension of Σ")
    mid = _group_method_id(method)
    opts = make_group_opts(**kwargs)
    info = GroupInfo()
    Vf = None if V is None else _f(V)
    _need(Vf is None or (Vf.ndim == 2 and Vf.shape[0] == p), f"V must have {p} rows")
    _need(float(m) >= 1, f"m should be 1 or larger but was {m}.")
    nv = 0 if Vf is None else Vf.shape[1]
    S = np.empty((p, p), order="F")
    obj = C.c_double()
    rc = ctx._lib.kb_solve_s_group(ctx.h, _ptr(Sigma), p, _ptr(groups), mid, float(m), C.byref(opts), _ptr(Vf), nv,
                                   _ptr(S), C.byref(obj), C.byref(info))
    check(ctx.h, rc)
    d = _group_info_dict(info, str(method).lstrip(":"))
    if opts.verbose:
        _print_group_trace(d)
    gam = np.array([info.gamma_equi]) if str(method).lstrip(":") == "equi" else np.zeros(0)
    out = (S, gam, obj.value)
    return out + (d,) if return_info else out


def solve_s_group_sharded(Sigma_block, groups, method, reduce, m=1, V=None, ctx: Optional[Context] = None, **kwargs):
    """One rank's part of ``solve_s_group`` on a block-diagonal Σ (kb_solve_s_group_sharded).
    ``reduce(op, values)`` must all-reduce the numpy array ``values`` IN PLACE across the ranks holding the other
    diagonal blocks (op: 0 = min, 1 = sum, 2 = max).  Returns (S_block, obj_local, info)."""
    ctx = ctx or default_context()
    Sigma_block = _f(Sigma_block)
    p = Sigma_block.shape[0]
    groups = np.ascontiguousarray(groups, dtype=np.int64)
    opts = make_group_opts(**kwargs)
    info = GroupInfo()
    Vf = None if V is None else _f(V)

    def _hook(user, op, vals, n):
        try:
            arr = np.ctypeslib.as_array(vals, shape=(n,))
            reduce(int(op), arr)
            return 0
        except Exception:                       # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return 1

    cb = _lib.REDUCE_FN(_hook)
    S = np.empty((p, p), order="F")
    obj = C.c_double()
    rc = ctx._lib.kb_solve_s_group_sharded(ctx.h, _ptr(Sigma_block), p, _ptr(groups), _group_method_id(method), float(m),
                                           C.byref(opts), _ptr(Vf), 0 if Vf is None else Vf.shape[1],
                                           C.cast(cb, C.c_void_p), None, _ptr(S), C.byref(obj), C.byref(info))
    check(ctx.h, rc)
    return S, obj.value, _group_info_dict(info, str(method).lstrip(":"))


@dataclass
class GaussianGroupKnockoff:
    """src/struct.jl:24-34."""
    X: np.ndarray
    Xko: np.ndarray
    groups: np.ndarray
    S: np.ndarray
    gammas: np.ndarray
    m: int
    Sigma: np.ndarray
    method: str
    obj: float
    info: dict = field(default_factory=dict)


def modelX_gaussian_group_knockoffs(X, method, groups, mu, Sigma, m=1, V=None, Z=None, seed=0,
                                    ctx: Optional[Context] = None, **kwargs) -> GaussianGroupKnockoff:
    """``modelX_gaussian_group_knockoffs(X, method, groups, μ, Σ; m, kwargs...)`` -- src/group.jl:109-127."""
    ctx = ctx or default_context()
    X = _f(X)
    n, p = X.shape
    groups = np.ascontiguousarray(groups, dtype=np.int64)
    if groups.shape != (p,):                                         # group.jl:119-120
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "Length of groups should be equal to number of variables")
    m = int(m)
    _need(m >= 1, f"m should be 1 or larger but was {m}.")
    Sigma = _f(Sigma)
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    _need(Sigma.shape == (p, p) and mu.shape == (p,), "dimension mismatch between X, mu and Sigma")
    mid = _group_method_id(method)
    opts = make_group_opts(**kwargs)
    info = GroupInfo()
    Vf = None if V is None else _f(V)
    _need(Vf is None or (Vf.ndim == 2 and Vf.shape[0] == p), f"V must have {p} rows")
    Zf = None if Z is None else _f(Z)
    if Zf is not None and Zf.shape != (n, m * p):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"Z must be {n} x {m * p}")
    Xko = np.empty((n, m * p), order="F")
    S = np.empty((p, p), order="F")
    obj = C.c_double()
    rc = ctx._lib.kb_modelx_gaussian_group_knockoffs(
        ctx.h, _ptr(X), n, p, _ptr(groups), _ptr(mu), _ptr(Sigma), mid, m, C.byref(opts), _ptr(Vf),
        0 if Vf is None else Vf.shape[1], _ptr(Zf), int(seed), 0, _ptr(Xko), _ptr(S), C.byref(obj), C.byref(info))
    check(ctx.h, rc)
    d = _group_info_dict(info, str(method).lstrip(":"))
    Ssym = np.triu(Sigma) + np.triu(Sigma, 1).T
    return GaussianGroupKnockoff(X, Xko, groups, S, np.zeros(0), m, Ssym, str(method).lstrip(":"), obj.value, d)



# ---- feature statistics and the knockoff filter ----------------------------------------
DEFAULT_FDRS = (0.01, 0.05, 0.1, 0.25, 0.5)                        # fit_lasso.jl:196


def _filter_flag(method) -> int:
    name = str(method).lstrip(":")
    if name == "knockoff":
        return 0
    if name == "knockoff_plus":
        return 1
    raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"method should be :knockoff or :knockoff_plus but was {name}.")


def threshold(w, q, method="knockoff_plus", rej_bounds=10000, ctx: Optional[Context] = None) -> float:
    """``threshold(w, q, method, rej_bounds)`` -- src/threshold.jl:19-31."""
    ctx = ctx or default_context()
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = C.c_double()
    check(ctx.h, ctx._lib.kb_threshold(ctx.h, _ptr(w), w.size, float(q), _filter_flag(method), int(rej_bounds),
                                       C.byref(out)))
    return out.value


def mk_threshold(tau, kappa, m, q, method="knockoff_plus", rej_bounds=10000, ctx: Optional[Context] = None) -> float:
    """``mk_threshold(τ, κ, m, q, method, rej_bounds)`` -- src/threshold.jl:55-75."""
    ctx = ctx or default_context()
    if str(method).lstrip(":") != "knockoff_plus":                 # threshold.jl:59
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "Multiple knockoffs needs to use :knockoff_plus filtering method")
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    kappa = np.ascontiguousarray(kappa, dtype=np.int64)
    if tau.size != kappa.size:                                     # threshold.jl:60
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "Length of τ and κ should be the same")
    out = C.c_double()
    check(ctx.h, ctx._lib.kb_mk_threshold(ctx.h, _ptr(tau), _ptr(kappa), tau.size, int(m), float(q), int(rej_bounds),
                                          C.byref(out)))
    return out.value


def MK_statistics(T0, Tk, ctx: Optional[Context] = None):
    """``MK_statistics(T0, Tk)`` -- src/threshold.jl:96-120 (``Tk`` a list of m > 1 vectors: returns (κ, τ, W))
    and :134-142 (``Tk`` one vector: returns W)."""
    ctx = ctx or default_context()
    T0 = np.ascontiguousarray(T0, dtype=np.float64)
    p = T0.size
    single = np.ndim(Tk[0]) == 0 if len(Tk) else True
    Tm = np.ascontiguousarray(Tk, dtype=np.float64).reshape(1 if single else len(Tk), -1)
    m = Tm.shape[0]
    if Tm.shape[1] != p:                                           # threshold.jl:102, :136
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "Length of T0 should equal all vectors in Tk")
    W = np.empty(p)
    kappa = np.zeros(p, dtype=np.int64)
    tau = np.zeros(p)
    check(ctx.h, ctx._lib.kb_mk_statistics(ctx.h, _ptr(T0), _ptr(Tm), p, m, _ptr(kappa), _ptr(tau), _ptr(W)))
    return W if single else (kappa, tau, W)


def marginal_stats(y, X, Xko, m=1, ctx: Optional[Context] = None):
    """Importance scores of ``fit_marginal`` (src/fit_lasso.jl:214-222): (T0, [T1, ..., Tm])."""
    ctx = ctx or default_context()
    X, Xko = _f(X), _f(Xko)
    n, p = X.shape
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    if y.size != n or Xko.shape != (n, m * p):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "dimension mismatch between y, X and Xko")
    T = np.empty((m + 1) * p)
    check(ctx.h, ctx._lib.kb_marginal_stats(ctx.h, _ptr(y), _ptr(X), _ptr(Xko), n, p, int(m), _ptr(T)))
    return T[:p], [T[(k + 1) * p:(k + 2) * p] for k in range(m)]


@dataclass
class MarginalKnockoffFilter:
    """src/struct.jl:134-146 (``d`` is always Normal here; ``kappa``/``tau``/``T0``/``Tk`` are extras)."""
    y: np.ndarray
    X: np.ndarray
    ko: object
    W: np.ndarray
    taus: np.ndarray
    m: int
    selected: list
    fdr_target: np.ndarray
    kappa: Optional[np.ndarray] = None
    tau: Optional[np.ndarray] = None
    T0: Optional[np.ndarray] = None
    Tk: Optional[list] = None


def _unpack_filter(p, m, nw, T, W, kappa, tau, taus, sel):
    W = W[:nw]
    selected = [np.flatnonzero(sel[f * nw:(f + 1) * nw]) + 1 for f in range(len(taus))]   # 1-based like findall
    return dict(W=W, taus=taus, selected=selected, kappa=kappa[:nw] if m > 1 else None, tau=tau[:nw] if m > 1 else None,
                T0=T[:p], Tk=[T[(k + 1) * p:(k + 2) * p] for k in range(m)])


def fit_marginal(y, X_or_ko, mu=None, Sigma=None, method="maxent", m=1, fdrs=DEFAULT_FDRS, groups=None,
                 filter_method="knockoff_plus", Z=None, seed=0, keep_knockoffs=True, ctx: Optional[Context] = None,
                 **kwargs) -> MarginalKnockoffFilter:
    """``fit_marginal`` -- src/fit_lasso.jl:188-257.

    ``fit_marginal(y, ko)`` (``ko`` a GaussianKnockoff / GaussianGroupKnockoff) scores an existing knockoff;
    ``fit_marginal(y, X, mu, Sigma, method=..., m=...)`` generates the knockoffs first.  Without ``groups`` that is
    ONE fused library call in which X̃ is scored chunk by chunk on the device (``keep_knockoffs=False``: X̃ is never
    copied back, ``ko.Xko`` is None); with ``groups`` the group knockoffs are generated, then scored.  Selected
    indices are 1-based like the reference's ``findall``."""
    ctx = ctx or default_context()
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    fd = np.ascontiguousarray(fdrs, dtype=np.float64)
    plus = _filter_flag(filter_method)
    if hasattr(X_or_ko, "Xko") and hasattr(X_or_ko, "X"):        # any knockoff struct (src/struct.jl)
        ko = X_or_ko
    elif groups is not None:
        if mu is None:
            Sigma, mu, _ = cov_shrink_lw(X_or_ko, ctx=ctx)
        ko = modelX_gaussian_group_knockoffs(X_or_ko, method, groups, mu, Sigma, m=m, Z=Z, seed=seed, ctx=ctx, **kwargs)
    else:
        X = _f(X_or_ko)
        n, p = X.shape
        m = int(m)
        if m < 1:
            raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"m should be 1 or larger but was {m}.")
        if y.size != n:
            raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "dimension mismatch between y and X")
        if Sigma is None:
            Sigma, mu, _ = cov_shrink_lw(X, ctx=ctx)               # fit_lasso.jl:199 -> modelX.jl:47-49
        Sigma = _f(Sigma)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        opts = make_opts(**kwargs)
        info = SolveInfo()
        Zf = None if Z is None else _f(Z)
        if Zf is not None and Zf.shape != (n, m * p):
            raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"Z must be {n} x {m * p}")
        Xko = np.empty((n, m * p), order="F") if keep_knockoffs else None
        s = np.empty(p)
        T, W = np.empty((m + 1) * p), np.empty(p)
        kappa, tau = np.zeros(p, dtype=np.int64), np.zeros(p)
        taus, sel = np.empty(fd.size), np.zeros(p * max(1, fd.size), dtype=np.int32)
        rc = ctx._lib.kb_fit_marginal(ctx.h, _ptr(y), _ptr(X), n, p, _ptr(mu), _ptr(Sigma), _method_id(method), m,
                                      C.byref(opts), _ptr(fd), fd.size, plus, _ptr(Zf), int(seed), _ptr(Xko), _ptr(s),
                                      _ptr(T), _ptr(W), _ptr(kappa), _ptr(tau), _ptr(taus), _ptr(sel), C.byref(info))
        check(ctx.h, rc)
        Ssym = np.triu(Sigma) + np.triu(Sigma, 1).T
        ko = GaussianKnockoff(X, Xko, s, Ssym, str(method).lstrip(":"), m, _info_dict(info, str(method).lstrip(":")))
        r = _unpack_filter(p, m, p, T, W, kappa, tau, taus, sel)
        return MarginalKnockoffFilter(y, X, ko, r["W"], r["taus"], m, r["selected"], fd, r["kappa"], r["tau"], r["T0"], r["Tk"])
    # score an existing knockoff (fit_lasso.jl:201-257)
    X, Xko, m = _f(ko.X), _f(ko.Xko), int(ko.m)
    n, p = X.shape
    if y.size != n:
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "dimension mismatch between y and X")
    g = getattr(ko, "groups", None)
    gi = None if g is None else np.ascontiguousarray(g, dtype=np.int64)
    T, W = np.empty((m + 1) * p), np.empty(p)
    kappa, tau = np.zeros(p, dtype=np.int64), np.zeros(p)
    taus, sel = np.empty(fd.size), np.zeros(p * max(1, fd.size), dtype=np.int32)
    nw = C.c_int64()
    rc = ctx._lib.kb_marginal_filter(ctx.h, _ptr(y), _ptr(X), _ptr(Xko), n, p, m, _ptr(gi), _ptr(fd), fd.size, plus,
                                     _ptr(T), _ptr(W), _ptr(kappa), _ptr(tau), _ptr(taus), _ptr(sel), C.byref(nw))
    check(ctx.h, rc)
    r = _unpack_filter(p, m, nw.value, T, W, kappa, tau, taus, sel)
    return MarginalKnockoffFilter(y, X, ko, r["W"], r["taus"], m, r["selected"], fd, r["kappa"], r["tau"], r["T0"], r["Tk"])



# ---- approximate (block-diagonal covariance) knockoffs --------------------------------
@dataclass
class ApproxGaussianKnockoff:
    """src/struct.jl:15-22.  ``Sigma`` is the dense block-diagonal matrix like the reference's (None with
    ``dense_sigma=False``); ``blocks`` holds the windows' Σ̂_w."""
    X: np.ndarray
    Xko: np.ndarray
    s: np.ndarray
    Sigma: Optional[np.ndarray]
    method: str
    m: int
    mu: Optional[np.ndarray] = None
    gamma: float = 1.0
    blocks: list = field(default_factory=list)
    window_ranges: list = field(default_factory=list)
    info: dict = field(default_factory=dict)


def _window_offsets(p, window_ranges=None, windowsize=500):
    """approx.jl:46-58 (``windowsize``) or explicit ranges: Julia ``a:b`` ranges / python (a, b) pairs, 1-based
    inclusive like the reference.  Returns the 0-based offsets array and the (a, b) list."""
    if window_ranges is None:
        if not windowsize > 1:                                     # approx.jl:45
            raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"windowsize should be > 1 but was {windowsize}")
        windows = -(-p // int(windowsize))
        window_ranges = [(w * windowsize + 1, p if w == windows - 1 else (w + 1) * windowsize) for w in range(windows)]
    rng = []
    for r in window_ranges:
        a, b = (r.start, r.stop - 1) if isinstance(r, range) else (int(r[0]), int(r[1]))
        rng.append((a, b))
    total = sum(b - a + 1 for a, b in rng)
    if total != p:                                                 # approx.jl:73-75
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"window_ranges have {total} dimensions but X has {p}")
    off = [0]
    for a, b in rng:
        if a != off[-1] + 1 or b < a:
            raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "window_ranges must tile 1:p in order (the reference "
                                                           "assembles BlockDiagonal(blocks) / vcat(s) in the given order)")
        off.append(b)
    return np.ascontiguousarray(off, dtype=np.int64), rng


def _blocks_from_packed(packed, off):
    out, o = [], 0
    for w in range(len(off) - 1):
        pw = int(off[w + 1] - off[w])
        out.append(packed[o:o + pw * pw].reshape((pw, pw), order="F"))
        o += pw * pw
    return out


def approx_modelX_gaussian_knockoffs(X, method, window_ranges=None, m=1, windowsize=500, Z=None, seed=0,
                                     dense_sigma=True, ctx: Optional[Context] = None, **kwargs) -> ApproxGaussianKnockoff:
    """``approx_modelX_gaussian_knockoffs(X, method[, window_ranges]; m, windowsize, kwargs...)`` -- src/approx.jl:37-102
    with the default covariance_approximator (Ledoit-Wolf linear shrinkage)."""
    ctx = ctx or default_context()
    X = _f(X)
    n, p = X.shape
    m = int(m)
    if m < 1:
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"m should be 1 or larger but was {m}.")
    off, rng = _window_offsets(p, window_ranges, windowsize)
    mid = _method_id(method)
    opts = make_opts(**kwargs)
    info = SolveInfo()
    Zf = None if Z is None else _f(Z)
    if Zf is not None and Zf.shape != (n, m * p):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"Z must be {n} x {m * p}")
    Xko = np.empty((n, m * p), order="F")
    s, mu = np.empty(p), np.empty(p)
    packed = np.empty(int(np.sum(np.diff(off) ** 2)))
    gamma = C.c_double()
    rc = ctx._lib.kb_approx_modelx_gaussian_knockoffs(ctx.h, _ptr(X), n, p, _ptr(off), len(off) - 1, mid, m, C.byref(opts),
                                                      _ptr(Zf), int(seed), _ptr(Xko), _ptr(s), _ptr(mu), _ptr(packed),
                                                      C.byref(gamma), C.byref(info))
    check(ctx.h, rc)
    blocks = _blocks_from_packed(packed, off)
    Sigma = None
    if dense_sigma:
        Sigma = np.zeros((p, p), order="F")
        for w, B in enumerate(blocks):
            Sigma[off[w]:off[w + 1], off[w]:off[w + 1]] = B
    return ApproxGaussianKnockoff(X, Xko, s, Sigma, str(method).lstrip(":"), m, mu, gamma.value, blocks, rng,
                                  _info_dict(info, str(method).lstrip(":")))


def feasible_gamma(Sigma, s, m=1, ctx: Optional[Context] = None) -> float:
    """``fzero(γ -> f(γ, s, Σ, m), 0, 1)`` of src/approx.jl:97, 105-109 on one dense matrix."""
    ctx = ctx or default_context()
    Sigma = _f(Sigma)
    s = np.ascontiguousarray(s, dtype=np.float64)
    p = _square(Sigma)
    _need(s.shape == (p,), f"s must have length {p}")
    out = C.c_double()
    check(ctx.h, ctx._lib.kb_feasible_gamma(ctx.h, _ptr(Sigma), p, _ptr(s), float(m), C.byref(out)))
    return out.value



# ---- representative group knockoffs ----------------------------------------------------
def choose_group_reps(Sigma, groups, threshold=0.5, prioritize_idx=None, ctx: Optional[Context] = None) -> np.ndarray:
    """``choose_group_reps(Σ::Symmetric, groups; threshold, prioritize_idx)`` -- src/group.jl:1929-2033.
    Indices (result and ``prioritize_idx``) are 1-based like the reference's."""
    ctx = ctx or default_context()
    Sigma = _f(Sigma)
    p = _square(Sigma)
    groups = np.ascontiguousarray(groups, dtype=np.int64)
    _need(groups.shape == (p,), "Length of groups should be equal to dimension of Σ")
    pr = None if prioritize_idx is None else np.ascontiguousarray(prioritize_idx, dtype=np.int64)
    out = np.zeros(p, dtype=np.int64)
    nr = C.c_int64()
    check(ctx.h, ctx._lib.kb_choose_group_reps(ctx.h, _ptr(Sigma), p, _ptr(groups), float(threshold), _ptr(pr),
                                               0 if pr is None else pr.size, _ptr(out), C.byref(nr)))
    return out[:nr.value].copy()


def cond_indep_corr(Sigma, groups, group_reps, ctx: Optional[Context] = None) -> np.ndarray:
    """``cond_indep_corr(Σ, groups, group_reps)`` -- src/group.jl:205-240 (1-based ``group_reps``)."""
    ctx = ctx or default_context()
    Sigma = _f(Sigma)
    p = _square(Sigma)
    groups = np.ascontiguousarray(groups, dtype=np.int64)
    reps = np.ascontiguousarray(group_reps, dtype=np.int64)
    _need(groups.shape == (p,), "Length of groups should be equal to dimension of Σ")
    _need(reps.ndim == 1 and (reps.size == 0 or (reps.min() >= 1 and reps.max() <= p)), f"group_reps must be indices in 1:{p}")
    out = np.empty((p, p), order="F")
    check(ctx.h, ctx._lib.kb_cond_indep_corr(ctx.h, _ptr(Sigma), p, _ptr(groups), _ptr(reps), reps.size, _ptr(out)))
    return out


def solve_s_graphical_group(Sigma, groups, group_reps, method, m=1, V=None, ctx: Optional[Context] = None,
                            return_info=False, **kwargs):
    """``solve_s_graphical_group(Σ::Symmetric, groups, group_reps, method; m, kwargs...)`` -- src/group.jl:268-303.
    Returns (S, D, obj)."""
    ctx = ctx or default_context()
    Sigma = _f(Sigma)
    p = _square(Sigma)
    groups = np.ascontiguousarray(groups, dtype=np.int64)
    reps = np.ascontiguousarray(group_reps, dtype=np.int64)
    _need(groups.shape == (p,), "Length of groups should be equal to dimension of Σ")
    _need(reps.ndim == 1 and reps.size >= 1 and reps.min() >= 1 and reps.max() <= p, f"group_reps must be indices in 1:{p}")
    r = reps.size
    opts = make_group_opts(**kwargs)
    info = GroupInfo()
    Vf = None if V is None else _f(V)
    S = np.empty((r, r), order="F")
    D = np.empty((p, p), order="F")
    obj = C.c_double()
    check(ctx.h, ctx._lib.kb_solve_s_graphical_group(ctx.h, _ptr(Sigma), p, _ptr(groups), _ptr(reps), r,
                                                     _group_method_id(method), float(m), C.byref(opts), _ptr(Vf),
                                                     0 if Vf is None else Vf.shape[1], _ptr(S), _ptr(D), C.byref(obj),
                                                     C.byref(info)))
    out = (S, D, obj.value)
    return out + (_group_info_dict(info, str(method).lstrip(":")),) if return_info else out


@dataclass
class GaussianRepGroupKnockoff:
    """src/struct.jl:36-48."""
    X: np.ndarray
    Xko: np.ndarray
    groups: np.ndarray
    group_reps: np.ndarray
    S11: np.ndarray
    S: np.ndarray
    m: int
    Sigma: np.ndarray
    method: str
    obj: float
    enforce_cond_indep: bool
    info: dict = field(default_factory=dict)


def modelX_gaussian_rep_group_knockoffs(X, method, groups, mu=None, Sigma=None, m=1, rep_threshold=0.5,
                                        enforce_cond_indep=False, group_reps=None, V=None, Z=None, seed=0,
                                        ctx: Optional[Context] = None, **kwargs) -> GaussianRepGroupKnockoff:
    """``modelX_gaussian_rep_group_knockoffs(X, method, groups[, μ, Σ]; m, rep_threshold, enforce_cond_indep, kwargs...)``
    -- src/group.jl:155-203.  ``group_reps`` (1-based) overrides choose_group_reps."""
    ctx = ctx or default_context()
    X = _f(X)
    n, p = X.shape
    groups = np.ascontiguousarray(groups, dtype=np.int64)
    if groups.shape != (p,):                                       # group.jl:183
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "Dimensions of X and groups doesn't match")
    if (mu is None) != (Sigma is None):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "give both mu and Sigma, or neither")
    if Sigma is None:
        Sigma, mu, _ = cov_shrink_lw(X, ctx=ctx)                   # group.jl:164-165
    Sigma = _f(Sigma)
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    m = int(m)
    if m < 1:
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"m should be 1 or larger but was {m}.")
    opts = make_group_opts(**kwargs)
    info = GroupInfo()
    Vf = None if V is None else _f(V)
    Zf = None if Z is None else _f(Z)
    if Zf is not None and Zf.shape != (n, m * p):
        raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"Z must be {n} x {m * p}")
    rin = None if group_reps is None else np.ascontiguousarray(group_reps, dtype=np.int64)
    Xko = np.empty((n, m * p), order="F")
    reps = np.zeros(p, dtype=np.int64)
    nr = C.c_int64()
    S11 = np.empty(p * p)
    D = np.empty((p, p), order="F")
    obj = C.c_double()
    rc = ctx._lib.kb_modelx_gaussian_rep_group_knockoffs(
        ctx.h, _ptr(X), n, p, _ptr(groups), _ptr(mu), _ptr(Sigma), _group_method_id(method), m, float(rep_threshold),
        int(bool(enforce_cond_indep)), C.byref(opts), _ptr(rin), 0 if rin is None else rin.size, _ptr(Vf),
        0 if Vf is None else Vf.shape[1], _ptr(Zf), int(seed), _ptr(Xko), _ptr(reps), C.byref(nr), _ptr(S11), _ptr(D),
        C.byref(obj), C.byref(info))
    check(ctx.h, rc)
    r = nr.value
    Ssym = np.triu(Sigma) + np.triu(Sigma, 1).T
    return GaussianRepGroupKnockoff(X, Xko, groups, reps[:r].copy(), S11[:r * r].reshape((r, r), order="F").copy(), D, m, Ssym,
                                    str(method).lstrip(":"), obj.value, bool(enforce_cond_indep),
                                    _group_info_dict(info, str(method).lstrip(":")))


# ---- hooks used by tests ---------------------------------------------------------------------------
def hook_dgemm(A, B, C_in=None, transa=False, transb=False, alpha=1.0, beta=0.0, ctx: Optional[Context] = None):
    ctx = ctx or default_context()
    A, B = _f(A), _f(B)
    M = A.shape[1] if transa else A.shape[0]
    K = A.shape[0] if transa else A.shape[1]
    N = B.shape[0] if transb else B.shape[1]
    Cm = np.zeros((M, N), order="F") if C_in is None else np.array(C_in, dtype=np.float64, order="F")
    check(ctx.h, ctx._lib.kb_test_dgemm(ctx.h, int(transa), int(transb), M, N, K, float(alpha), _ptr(A), A.shape[0],
                                        _ptr(B), B.shape[0], float(beta), _ptr(Cm), M))
    return Cm


def hook_factor_band(Lblocks, ctx: Optional[Context] = None) -> int:
    """Structure of factor blocks as the sampling stage sees it: 0 / 16..64 (two triangular products per copy) or -1."""
    ctx = ctx or default_context()
    Lb = _f(Lblocks)
    out = C.c_int32()
    check(ctx.h, ctx._lib.kb_test_factor_band(ctx.h, _ptr(Lb), Lb.shape[0], (Lb.shape[2] + 1) // 2, C.byref(out)))
    return out.value


def hook_potrf(A, ctx: Optional[Context] = None):
    ctx = ctx or default_context()
    A = _f(A)
    L = np.empty_like(A, order="F")
    check(ctx.h, ctx._lib.kb_test_potrf(ctx.h, _ptr(A), A.shape[0], _ptr(L)))
    return L


def hook_potrf_positive(A, ctx: Optional[Context] = None):
    """(L, d, modified) of the restated cholesky(Positive, Symmetric(A)) (csrc/positive.cu)."""
    ctx = ctx or default_context()
    A = _f(A)
    p = A.shape[0]
    L = np.empty((p, p), order="F")
    d = np.empty(p)
    mod = C.c_int32()
    check(ctx.h, ctx._lib.kb_test_potrf_positive(ctx.h, _ptr(A), p, _ptr(L), _ptr(d), C.byref(mod)))
    return L, d, mod.value


def hook_spd_inverse(A, ctx: Optional[Context] = None):
    ctx = ctx or default_context()
    A = _f(A)
    O = np.empty_like(A, order="F")
    check(ctx.h, ctx._lib.kb_test_spd_inverse(ctx.h, _ptr(A), A.shape[0], _ptr(O)))
    return O


def hook_randn(seed, row_offset, n, ncols, ctx: Optional[Context] = None):
    ctx = ctx or default_context()
    Z = np.empty((n, ncols), order="F")
    check(ctx.h, ctx._lib.kb_test_randn(ctx.h, int(seed), int(row_offset), int(n), int(ncols), _ptr(Z)))
    return Z


# ---- several GPUs behind one handle (kb_create_multi) -----------------------------------
class MultiContext:
    """``kb_create_multi(devices)``: one host process, one context + one host thread per listed device.  Rows of X are
    sharded for the Gram (one peer-to-peer all-reduce over NVLink) and for ``condition``; independent solves and LD
    blocks of a block-diagonal group problem are dealt to the devices; a single CCD solve stays on one device."""

    def __init__(self, devices):
        self._lib = _lib.load()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self._lib.kb_create_multi(C.byref(h), devs, len(devices))
        if rc != 0:
            raise _lib.KnockoffsError(rc, f"kb_create_multi({list(devices)}) failed (knockoffs_b200 has no CPU fallback)")
        self.h = h
        self.devices = [int(d) for d in devices]

    def close(self):
        if getattr(self, "h", None):
            self._lib.kb_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = (self._lib.kb_multi_last_error(self.h) or b"").decode(errors="replace")
            if rc == _lib.KB_ERR_NOT_PD:
                raise _lib.PosDefException(rc, msg)
            raise _lib.KnockoffsError(rc, msg)

    @property
    def launches(self) -> int:
        return sum(int(self._lib.kb_launch_count(self._lib.kb_multi_ctx(self.h, i))) for i in range(len(self.devices)))

    def gram(self, X, center=True, denom=None):
        X = _f(X)
        n, p = X.shape
        G = np.empty((p, p), order="F")
        mu = np.empty(p)
        self._check(self._lib.kb_multi_gram(self.h, _ptr(X), n, p, int(bool(center)), float(denom or 0.0), _ptr(G), _ptr(mu)))
        return G, mu

    def condition(self, X, mu, Sigma, S, m=1, Z=None, seed=0):
        X = _f(X)
        n, p = X.shape
        Sigma = _f(Sigma)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        S = np.asarray(S, dtype=np.float64)
        is_diag = S.ndim == 1
        S = np.ascontiguousarray(S) if is_diag else _f(S)
        if Sigma.shape != (p, p) or mu.shape != (p,) or S.shape != ((p,) if is_diag else (p, p)):
            raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "dimension mismatch between X, mu, Sigma and S")
        Zf = None if Z is None else _f(Z)
        if Zf is not None and Zf.shape != (n, int(m) * p):
            raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, f"Z must be {n} x {int(m) * p}")
        Xko = np.empty((n, int(m) * p), order="F")
        self._check(self._lib.kb_multi_condition(self.h, _ptr(X), n, p, _ptr(mu), _ptr(Sigma), _ptr(S), int(is_diag), int(m),
                                                 _ptr(Zf), int(seed), _ptr(Xko)))
        return Xko

    def modelX_gaussian_knockoffs(self, X, method, mu, Sigma, m=1, Z=None, seed=0, s_init=None, **kwargs) -> "GaussianKnockoff":
        X = _f(X)
        n, p = X.shape
        Sigma = _f(Sigma)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        if Sigma.shape != (p, p) or mu.shape != (p,):
            raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "dimension mismatch between X, mu and Sigma")
        opts = make_opts(**kwargs)
        info = SolveInfo()
        Zf = None if Z is None else _f(Z)
        si = None if s_init is None else np.ascontiguousarray(s_init, dtype=np.float64)
        if si is not None and si.shape != (p,):
            raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "s_init must have length p")
        Xko = np.empty((n, int(m) * p), order="F")
        s = np.empty(p)
        self._check(self._lib.kb_multi_modelx_gaussian_knockoffs(self.h, _ptr(X), n, p, _ptr(mu), _ptr(Sigma), _method_id(method),
                                                                 int(m), C.byref(opts), _ptr(si), _ptr(Zf), int(seed), _ptr(Xko),
                                                                 _ptr(s), C.byref(info)))
        Ssym = np.triu(Sigma) + np.triu(Sigma, 1).T
        return GaussianKnockoff(X, Xko, s, Ssym, str(method).lstrip(":"), int(m), _info_dict(info, str(method).lstrip(":")))

    def solve_s_many(self, Sigma, methods, ms, **kwargs):
        """Independent ``solve_s(Σ, methods[q]; m = ms[q])`` problems dealt round-robin to the devices ("replicas only")."""
        Sigma = _f(Sigma)
        p = Sigma.shape[0]
        nprob = len(methods)
        mids = np.array([_method_id(x) for x in methods], dtype=np.int32)
        mf = np.array([float(x) for x in ms], dtype=np.float64)
        opts = make_opts(**kwargs)
        infos = (SolveInfo * nprob)()
        s = np.empty((p, nprob), order="F")
        self._check(self._lib.kb_multi_solve_s_many(self.h, _ptr(Sigma), p, nprob, _ptr(mids), _ptr(mf), C.byref(opts), _ptr(s),
                                                    C.cast(infos, C.c_void_p)))
        return [s[:, q].copy() for q in range(nprob)], [_info_dict(infos[q], str(methods[q]).lstrip(":")) for q in range(nprob)]

    def solve_s_group_blockdiag(self, Sigma_blocks, groups_blocks, method, m=1, **kwargs):
        """``solve_s_group`` on Σ = blockdiag(Sigma_blocks): block b on device b mod ndev, lockstep sweeps, the reference's
        global decisions reduced in-process.  Returns (list of S blocks, summed objective, list of info dicts)."""
        nb = len(Sigma_blocks)
        Sb = [_f(x) for x in Sigma_blocks]
        gb = [np.ascontiguousarray(g, dtype=np.int64) for g in groups_blocks]
        for x, g in zip(Sb, gb):
            if x.ndim != 2 or x.shape[0] != x.shape[1] or g.shape != (x.shape[0],):
                raise _lib.KnockoffsError(_lib.KB_ERR_INVALID, "every block needs a square Sigma and one group id per variable")
        pb = np.array([x.shape[0] for x in Sb], dtype=np.int64)
        out = [np.empty((x.shape[0], x.shape[0]), order="F") for x in Sb]
        arr = lambda lst: (C.c_void_p * nb)(*[x.ctypes.data for x in lst])       # noqa: E731
        opts = make_group_opts(**kwargs)
        infos = (GroupInfo * nb)()
        obj = C.c_double()
        self._check(self._lib.kb_multi_solve_s_group_blockdiag(self.h, nb, arr(Sb), _ptr(pb), arr(gb), _group_method_id(method),
                                                               float(m), C.byref(opts), arr(out), C.byref(obj),
                                                               C.cast(infos, C.c_void_p)))
        name = str(method).lstrip(":")
        return out, obj.value, [_group_info_dict(infos[b], name) for b in range(nb)]
