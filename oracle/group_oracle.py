"""CPU oracle (numpy/scipy) for the GROUP knockoff solvers of Knockoffs.jl v2.0.2 (src/group.jl).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as knockoffs_oracle.py).

Literal restatement, keeping the reference's formulation (two Cholesky factors L of (m+1)/m Σ − S and
C of S, triangular solves, Givens rank-1 up/downdates):
    solve_s_group                     src/group.jl:348-422
    initialize_S / solve_group_equi   :432-443, :479-495 (inverse_mat_sqrt :448-454)
    solve_group_{max_entropy,sdp,mvr}_hybrid   :870-919, :951-1025, :1057-1106
    _sdp/_mvr/_maxent_ccd_iter!       :1108-1215, :1217-1343, :1345-1458
    _maxent/_mvr/_sdp_pca_ccd_iter!   :1460-1523, :1525-1593, :1595-1673
    objectives                        :14-45, :1675-1727
    rank1/rank2_cholesky_update!      :1729-1768
    get_PCA_vectors                   :2161-2187

Third-party pieces that are not under /root/reference (parity UNPINNED against real Julia for these):
  * Optim.optimize(f, lb, ub, Brent(); abs_tol=1e-4) (group.jl:1303,1419,1624) -- restated below from the
    published algorithm (Brent 1973 / Optim.jl univariate/solvers/brent.jl), rel_tol = sqrt(eps), 1000 iterations.
  * eigen(BlockDiagonal) inside get_PCA_vectors -- the ORDER of the directions is implementation defined
    (SURVEY Q9); here: block by block, eigenvalues ascending inside a block, then the unit vectors; callers
    that need a specific order pass V explicitly (the GPU library takes the same V).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla

from .knockoffs_oracle import lowrankdowndate_turbo, lowrankupdate_turbo, solve_s


# ------------------------------------------------------------------------------------------------
# Optim.jl Brent (univariate)
# ------------------------------------------------------------------------------------------------
def brent(f, x_lower, x_upper, abs_tol=1e-4, rel_tol=math.sqrt(np.finfo(float).eps), iterations=1000):
    golden = 0.5 * (3.0 - math.sqrt(5.0))
    x = x_lower + golden * (x_upper - x_lower)
    fx = f(x)
    step = old_step = 0.0
    x1 = x2 = x
    f1 = f2 = fx
    it = 0
    while it < iterations:
        p = q = 0.0
        x_tol = rel_tol * abs(x) + abs_tol
        mid = (x_upper + x_lower) / 2
        if abs(x - mid) <= 2 * x_tol - (x_upper - x_lower) / 2:
            break
        it += 1
        if abs(old_step) > x_tol:
            r = (x - x1) * (fx - f2)
            q = (x - x2) * (fx - f1)
            p = (x - x2) * q - (x - x1) * r
            q = 2 * (q - r)
            if q > 0:
                p = -p
            else:
                q = -q
        if abs(p) < abs(q * old_step / 2) and p < q * (x_upper - x) and p < q * (x - x_lower):
            old_step = step
            step = p / q
            xt = x + step
            if (xt - x_lower) < 2 * x_tol or (x_upper - xt) < 2 * x_tol:
                step = x_tol if x < mid else -x_tol
        else:
            old_step = (x_upper - x) if x < mid else (x_lower - x)
            step = golden * old_step
        if abs(step) >= x_tol:
            xn = x + step
        else:
            xn = x + (x_tol if step > 0 else -x_tol)
        fn = f(xn)
        if fn < fx:
            if xn < x:
                x_upper = x
            else:
                x_lower = x
            x2, f2 = x1, f1
            x1, f1 = x, fx
            x, fx = xn, fn
        else:
            if xn < x:
                x_lower = xn
            else:
                x_upper = xn
            if fn <= f1 or x1 == x:
                x2, f2 = x1, f1
                x1, f1 = xn, fn
            elif fn <= f2 or x2 == x or x2 == x1:
                x2, f2 = xn, fn
    return x, fx


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def _group_index_lists(groups):
    """unique(groups) in order of first appearance, with member indices (findall)."""
    order, idx = [], {}
    for i, g in enumerate(groups):
        if g not in idx:
            idx[g] = []
            order.append(g)
        idx[g].append(i)
    return [np.array(idx[g]) for g in order]


def inverse_mat_sqrt(A, tol=1e-6):          # group.jl:448-454
    lam, phi = np.linalg.eigh(A)
    lam = np.where(lam < tol, tol, lam)
    return (phi / np.sqrt(lam)) @ phi.T


def sdp_block_objective(Sg_sigma, Sg):      # group.jl:35-45
    g = Sg.shape[0]
    return np.sum(np.abs(Sg_sigma - Sg)) / g ** 2


def group_block_objective(Sigma, S, groups, m, method):   # group.jl:14-32
    p = Sigma.shape[0]
    kap = (m + 1) / m
    if "sdp" in method or "equi" in method:
        return sum(sdp_block_objective(Sigma[np.ix_(ix, ix)], S[np.ix_(ix, ix)]) for ix in _group_index_lists(groups))
    if "maxent" in method:
        return np.linalg.slogdet(kap * Sigma - S + 1e-8 * np.eye(p))[1] + m * np.linalg.slogdet(S + 1e-8 * np.eye(p))[1]
    if "mvr" in method:
        return m ** 2 * np.trace(np.linalg.inv(S)) + np.trace(np.linalg.inv(kap * Sigma - S))
    raise ValueError("unrecognized method")


def solve_group_equi(Sigma, groups, m=1):   # group.jl:479-495
    p = Sigma.shape[0]
    Db = np.zeros((p, p))
    Sb = np.zeros((p, p))
    for ix in _group_index_lists(groups):
        blk = Sigma[np.ix_(ix, ix)]
        Db[np.ix_(ix, ix)] = inverse_mat_sqrt(blk)
        Sb[np.ix_(ix, ix)] = blk
    T = Db @ Sigma @ Db
    lam = np.linalg.eigvalsh((T + T.T) / 2).min()
    gamma = min(1.0, (m + 1) / m * lam)
    S = gamma * Sb
    return S, np.array([gamma]), group_block_objective(Sigma, S, groups, m, "equi")


def _chol_upper(A):
    return np.ascontiguousarray(sla.cholesky(A, lower=False, check_finite=False))


def initialize_S(Sigma, groups, m, eps=1e-8):   # group.jl:432-443
    S, _, _ = solve_group_equi(Sigma, groups, m)
    ev, evec = np.linalg.eigh(S)
    ev = np.where(ev < eps, eps, ev)
    S = (evec * ev) @ evec.T
    S = S / 2
    L = _chol_upper((m + 1) / m * Sigma - S)
    C = _chol_upper(S)
    return S, L, C


def get_PCA_vectors(Sigma, groups):         # group.jl:2161-2187 (direction ORDER: see module docstring)
    p = Sigma.shape[0]
    cols = []
    for ix in _group_index_lists(groups):
        _, vec = np.linalg.eigh(Sigma[np.ix_(ix, ix)])
        for k in range(len(ix)):
            v = np.zeros(p)
            v[ix] = vec[:, k]
            cols.append(v)
    for i in range(p):
        v = np.zeros(p)
        v[i] = 1.0
        cols.append(v)
    out, seen = [], set()
    for v in cols:                           # unique(..., dims=2)
        key = v.tobytes()
        if key not in seen:
            seen.add(key)
            out.append(v)
    return np.asfortranarray(np.stack(out, axis=1))


def _ldivT(U, b):     # ldiv!(x, UpperTriangular(F.factors)', b)
    return sla.solve_triangular(U, b, trans="T", lower=False, check_finite=False)


def _fb(U, b):        # forward_backward!: solves U'U x = b
    return sla.solve_triangular(U, _ldivT(U, b), trans="N", lower=False, check_finite=False)


def _update(U, v, plus):
    (lowrankupdate_turbo if plus else lowrankdowndate_turbo)(U, v.copy())


def rank1_cholesky_update(L, C, j, d):      # group.jl:1729-1741
    p = L.shape[0]
    e = np.zeros(p)
    e[j] = math.sqrt(abs(d))
    if d > 0:
        _update(L, e, False)
        _update(C, e, True)
    else:
        _update(L, e, True)
        _update(C, e, False)


def rank2_cholesky_update(L, C, i, j, d):   # group.jl:1743-1768
    p = L.shape[0]
    s1 = np.zeros(p)
    s2 = np.zeros(p)
    r = math.sqrt(abs(d / 2))
    s1[j] = s1[i] = s2[j] = r
    s2[i] = -r
    if d > 0:
        _update(L, s1, False); _update(L, s2, True)
        _update(C, s1, True); _update(C, s2, False)
    else:
        _update(L, s1, True); _update(L, s2, False)
        _update(C, s1, False); _update(C, s2, True)


def _pca_factor_update(L, C, v, d):         # group.jl:1494-1506 and twins
    u = math.sqrt(abs(d)) * v
    if d > 0:
        _update(L, u, False)
        _update(C, u, True)
    else:
        _update(L, u, True)
        _update(C, u, False)


def offdiag_maxent_obj(d, m, aij, aii, ajj, bij, bii, bjj):      # group.jl:1695-1700 (incl. quirk Q5)
    in1 = (1 - d * aij) ** 2 - d * d * aii * ajj
    in2 = (1 + d * bij) ** 2 - d * d * bjj * bii
    if in1 <= 0:
        return float("nan")          # Julia: falls through to log(in1 <= 0) -> DomainError
    if in2 <= 0:
        return -float("inf")
    return -math.log(in1) - m * math.log(in2)


def offdiag_mvr_obj(d, m, aij, aii, ajj, bij, bii, bjj, cij, cii, cjj, dij, dii, djj):   # group.jl:1701-1707
    denom1 = (1 + d * bij) ** 2 - d * d * bii * bjj
    denom2 = (1 - d * aij) ** 2 - d * d * aii * ajj
    numer1 = -m ** 2 * d * ((cij * bij - cjj * bii - cii * bjj + cij * bij) * d + 2 * cij)
    numer2 = d * ((-dij * aij + djj * aii + dii * ajj - dij * aij) * d + 2 * dij)
    return numer1 / denom1 + numer2 / denom2


def diag_mvr_obj_root(m, ajj, bjj, cjj, djj):      # group.jl:1708-1716
    a = -ajj ** 2 * m ** 2 * cjj + bjj ** 2 * djj
    b = 2 * ajj * m ** 2 * cjj + 2 * bjj * djj
    c = djj - m ** 2 * cjj
    if a == 0 and c == 0:
        return 0.0, 0.0
    with np.errstate(all="ignore"):
        disc = np.sqrt(np.float64(b * b - 4 * a * c))
        return float((-b + disc) / np.float64(2 * a)), float((-b - disc) / np.float64(2 * a))


def pca_sdp_obj(d, Sg_sigma, Sg, v):        # group.jl:1717-1727
    g = len(v)
    return np.sum(np.abs(Sg_sigma - Sg - d * np.outer(v, v))) / g ** 2


def _bad(x):
    return abs(x) < 1e-15 or math.isnan(x) or math.isinf(x)


def _offdiag_bounds(aij, aii, ajj, bij, bii, bjj, eps, sdp_style):
    with np.errstate(all="ignore"):
        s1 = (aij - math.sqrt(aii * ajj)) / (aij ** 2 - aii * ajj)
        s2 = (aij + math.sqrt(aii * ajj)) / (aij ** 2 - aii * ajj)
        d1 = (-bij - math.sqrt(bii * bjj)) / (bij ** 2 - bii * bjj)
        d2 = (-bij + math.sqrt(bii * bjj)) / (bij ** 2 - bii * bjj)
    if s1 > s2:
        s1, s2 = s2, s1
    if d1 > d2:
        d1, d2 = d2, d1
    if sdp_style:                    # group.jl:1176-1177 (epsilon outside max/min: quirk Q4)
        lb = max(s1, d1, -2 / (bii + 2 * bij + bjj)) + eps
        ub = min(s2, d2, 2 / (aii + 2 * aij + ajj)) - eps
    else:                            # group.jl:1299-1300, 1415-1416
        lb = max(s1, d1, -2 / (bii + 2 * bij + bjj) + eps)
        ub = min(s2, d2, 2 / (aii + 2 * aij + ajj) - eps)
    return lb, ub


# ------------------------------------------------------------------------------------------------
# inner iterations
# ------------------------------------------------------------------------------------------------
def _pca_iter(method, S, L, C, V, Sigma, obj, m, niter, tol, eps, sdp_state, trace):
    p = S.shape[0]
    converged = niter == 0
    for _ in range(niter):
        max_delta = 0.0
        obj_new = obj
        for jv in range(V.shape[1]):
            v = V[:, jv]
            w = _ldivT(L, v)
            u = _ldivT(C, v)
            b = float(u @ u)        # vt_Sinv_v
            a = float(w @ w)        # vt_Dinv_v
            if method == "mvr":
                w2 = _fb(L, v)
                u2 = _fb(C, v)
                dd = float(w2 @ w2)
                cc = float(u2 @ u2)
            lb = -1 / b + eps
            ub = 1 / a - eps
            if lb >= ub:
                continue
            if method == "maxent":
                d = (m * b - a) / ((m + 1) * b * a)
                d = min(max(d, lb), ub)
                with np.errstate(all="ignore"):
                    change = float(np.log(np.float64(1 - d * a)) + m * np.log(np.float64(1 + d * b)))
                if change < 0 or _bad(d):
                    continue
            elif method == "mvr":
                x1, x2 = diag_mvr_obj_root(m, a, b, cc, dd)
                d = x1 if lb < x1 < ub else (x2 if lb < x2 < ub else float("nan"))
                change = -m ** 2 * d * cc / (1 + d * b) + d * dd / (1 - d * a)
                if change > 0 or _bad(d):
                    continue
            else:   # sdp
                gi = sdp_state["v_groups"][jv]
                ix = sdp_state["group_idx"][gi]
                Sg_sigma = Sigma[np.ix_(ix, ix)]
                Sg = S[np.ix_(ix, ix)]
                vg = v[ix]
                xmin, fmin = brent(lambda t: pca_sdp_obj(t, Sg_sigma, Sg, vg), lb, ub)
                d = min(max(xmin, lb), ub)
                change = fmin - sdp_state["group_obj"][gi]
                if _bad(d) or change > 0.01:
                    continue
                sdp_state["group_obj"][gi] = fmin
            S += d * np.outer(v, v)
            obj_new += change
            _pca_factor_update(L, C, v, d)
            max_delta = max(max_delta, abs(d))
        trace.append(("pca", obj_new, max_delta))
        change_obj = abs((obj_new - obj) / obj)
        obj = obj_new
        if change_obj < tol or max_delta < 1e-4:
            converged = True
            break
    return converged, obj


def _ccd_iter(method, S, L, C, Sigma, obj, m, group_sizes, niter, tol, eps, trace):
    p = S.shape[0]
    converged = niter == 0
    for _ in range(niter):
        max_delta = 0.0
        obj_new = obj
        offset = 0
        for gs in group_sizes:
            for idx in range(gs):                                  # diagonal entries
                j = idx + offset
                ej = np.zeros(p)
                ej[j] = 1
                u = _ldivT(L, ej)
                v = _ldivT(C, ej)
                ajj, bjj = float(u @ u), float(v @ v)
                ub = 1 / ajj - eps
                lb = -1 / bjj + eps
                if method == "mvr":
                    vv = _fb(C, ej)
                    uu = _fb(L, ej)
                    cjj, djj = float(vv @ vv), float(uu @ uu)
                if lb >= ub:
                    continue
                if method == "sdp":
                    d = min(max(Sigma[j, j] - S[j, j], lb), ub)
                    change = (abs(Sigma[j, j] - S[j, j] - d) - abs(Sigma[j, j] - S[j, j])) / gs ** 2
                    if _bad(d) or change > 0.01:
                        continue
                elif method == "maxent":
                    d = (m * bjj - ajj) / ((m + 1) * ajj * bjj)
                    d = min(max(d, lb), ub)
                    with np.errstate(all="ignore"):
                        change = float(np.log(np.float64(1 - d * ajj)) + m * np.log(np.float64(1 + d * bjj)))
                    if change < 0 or _bad(d):
                        continue
                else:
                    x1, x2 = diag_mvr_obj_root(m, ajj, bjj, cjj, djj)
                    d = x1 if lb < x1 < ub else (x2 if lb < x2 < ub else float("nan"))
                    change = -m ** 2 * d * cjj / (1 + d * bjj) + d * djj / (1 - d * ajj)
                    if change > 0 or _bad(d):
                        continue
                S[j, j] += d
                obj_new += change
                rank1_cholesky_update(L, C, j, d)
                max_delta = max(max_delta, abs(d))
            for idx1 in range(gs):                                 # off-diagonal entries
                for idx2 in range(idx1 + 1, gs):
                    i, j = idx2 + offset, idx1 + offset
                    ei = np.zeros(p); ej = np.zeros(p)
                    ei[i] = 1; ej[j] = 1
                    u = _ldivT(L, ei); v = _ldivT(L, ej)
                    aij, aii, ajj = float(u @ v), float(u @ u), float(v @ v)
                    u = _ldivT(C, ei); v = _ldivT(C, ej)
                    bij, bii, bjj = float(u @ v), float(u @ u), float(v @ v)
                    if method == "mvr":
                        u = _fb(C, ei); v = _fb(C, ej)
                        cij, cii, cjj = float(u @ v), float(u @ u), float(v @ v)
                        u = _fb(L, ei); v = _fb(L, ej)
                        dij, dii, djj = float(u @ v), float(u @ u), float(v @ v)
                    lb, ub = _offdiag_bounds(aij, aii, ajj, bij, bii, bjj, eps, method == "sdp")
                    if lb >= ub:
                        continue
                    if method == "sdp":
                        d = min(max(Sigma[i, j] - S[i, j], lb), ub)
                        change = (2 * abs(Sigma[i, j] - S[i, j] - d) - 2 * abs(Sigma[i, j] - S[i, j])) / gs ** 2
                        if _bad(d) or change > 0.01:
                            continue
                    elif method == "maxent":
                        xmin, fmin = brent(lambda t: offdiag_maxent_obj(t, m, aij, aii, ajj, bij, bii, bjj), lb, ub)
                        d = min(max(xmin, lb), ub)
                        change = -fmin
                        if change < 0 or _bad(d):
                            continue
                    else:
                        xmin, fmin = brent(lambda t: offdiag_mvr_obj(t, m, aij, aii, ajj, bij, bii, bjj,
                                                                     cij, cii, cjj, dij, dii, djj), lb, ub)
                        d = min(max(xmin, lb), ub)
                        change = fmin
                        if change > 0 or _bad(d):
                            continue
                    S[i, j] += d
                    S[j, i] += d
                    obj_new += change
                    rank2_cholesky_update(L, C, i, j, d)
                    max_delta = max(max_delta, abs(d))
            offset += gs
        trace.append(("ccd", obj_new, max_delta))
        change_obj = abs((obj_new - obj) / obj)
        obj = obj_new
        if change_obj < tol or max_delta < 1e-4:
            converged = True
            break
    return converged, obj


def solve_group_hybrid(Sigma, groups, method, m=1, outer_iter=100, inner_pca_iter=1, inner_ccd_iter=1, tol=1e-4,
                       eps=1e-6, V=None, trace=None):
    """solve_group_{max_entropy,sdp,mvr}_hybrid for CONTIGUOUS groups on a correlation matrix.
    Returns (S, obj, trace)."""
    Sigma = np.array(Sigma, dtype=float)
    groups = list(groups)
    idx_lists = _group_index_lists(groups)
    group_sizes = [len(ix) for ix in idx_lists]
    if V is None:
        V = get_PCA_vectors(Sigma, groups)
    S, L, C = initialize_S(Sigma, groups, m)
    trace = [] if trace is None else trace
    sdp_state = None
    if method == "maxent":
        obj = 2 * np.sum(np.log(np.diag(L))) + m * 2 * np.sum(np.log(np.diag(C)))     # group_maxent_obj :1678-1680
    elif method == "mvr":
        obj = group_block_objective(Sigma, S, groups, m, "mvr")
    else:
        gobj = [sdp_block_objective(Sigma[np.ix_(ix, ix)], S[np.ix_(ix, ix)]) for ix in idx_lists]
        obj = float(sum(gobj))
        if obj < eps:
            return S, obj, trace
        v_groups = []
        for jv in range(V.shape[1]):
            nz = int(np.flatnonzero(V[:, jv])[0])
            v_groups.append(next(gi for gi, ix in enumerate(idx_lists) if nz in ix))
        sdp_state = dict(group_idx=idx_lists, v_groups=v_groups, group_obj=gobj)
    for _ in range(outer_iter):
        c1, obj = _pca_iter(method, S, L, C, V, Sigma, obj, m, inner_pca_iter, tol, eps, sdp_state, trace)
        c2, obj = _ccd_iter(method, S, L, C, Sigma, obj, m, group_sizes, inner_ccd_iter, tol, eps, trace)
        if method == "sdp" and inner_pca_iter > 0:
            for gi, ix in enumerate(idx_lists):
                sdp_state["group_obj"][gi] = sdp_block_objective(Sigma[np.ix_(ix, ix)], S[np.ix_(ix, ix)])
        if c1 and c2:
            break
    return S, float(obj), trace


def solve_s_group(Sigma, groups, method, m=1, V=None, **kw):
    """src/group.jl:348-422.  Returns (S, gammas, obj).  V (optional, p x nv) is given in the CALLER's variable
    order and is row-permuted together with Σ (the reference computes it internally after permuting)."""
    Sigma = np.triu(Sigma) + np.triu(Sigma, 1).T
    groups = np.asarray(groups)
    p = Sigma.shape[0]
    if len(groups) != p:
        raise ValueError("Length of groups should be equal to dimension of Sigma")
    method = str(method).lstrip(":")
    sig = np.sqrt(np.diag(Sigma))
    iscor = bool(np.all(np.isclose(sig, 1.0, rtol=np.sqrt(np.finfo(float).eps), atol=0)))
    Scor = Sigma / np.outer(sig, sig)
    np.fill_diagonal(Scor, 1.0)
    perm = np.argsort(groups, kind="stable")                # sortperm
    gperm = groups.copy()
    permuted = False
    if not np.all(groups[:-1] <= groups[1:]):
        gperm = groups[perm]
        Scor = Scor[np.ix_(perm, perm)]
        permuted = True
    if len(np.unique(groups)) == p:
        s = solve_s(Scor, method, m=m, **kw)
        S, gam, obj = np.diag(s), np.zeros(0), 0.0
    elif method == "equi":
        S, gam, obj = solve_group_equi(Scor, list(gperm), m)
    elif method in ("maxent", "mvr", "sdp"):
        Vp = None if V is None else (np.asarray(V)[perm, :] if permuted else np.asarray(V))
        S, obj, _ = solve_group_hybrid(Scor, list(gperm), method, m=m, V=Vp, **kw)
        gam = np.zeros(0)
    else:
        raise ValueError(f"Method must be one of the group knockoff methods but was {method}")
    if permuted:
        iperm = np.argsort(perm, kind="stable")
        S = S[np.ix_(iperm, iperm)]
    if not iscor:
        S = S * np.outer(sig, sig)                          # cor2cov!
    return S, gam, obj
