"""CPU oracle (numpy, literal) for representative group knockoffs.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (only tests/, smoke() and bench.py's CPU legs may import it).

Restates, line by line (paths relative to /root/reference):
  src/group.jl:1929-2033  choose_group_reps (incl. the dependent-column pre-pass and prioritize_idx)
  src/group.jl:2050-2056  prioritize_variants         src/group.jl:2070-2098  select_one / select_best_rss_subset
  src/group.jl:205-240    cond_indep_corr             src/group.jl:268-303    solve_s_graphical_group
  src/group.jl:170-203    modelX_gaussian_rep_group_knockoffs
Pinned against the reference's own known answers for this path: test/runtests.jl:733-744 (issue 71: the 6 x 6
correlation matrix with groups [1,2,3,4,3,3] -> representatives [1,2,3,4] at threshold 0.5 and [1,2,3,4,5] at 0.999)
= tests/golden/reference_known_answers.json key K9; the nesting property of runtests.jl:712-716 is checked too.
Indices are 0-based here (the golden file keeps the reference's 1-based values).
"""
from __future__ import annotations

import numpy as np

from . import group_oracle as go
from . import knockoffs_oracle as ko


def _isapprox(a, b):
    """Julia isapprox(a, b): |a − b| <= sqrt(eps) * max(|a|, |b|)."""
    return abs(a - b) <= np.sqrt(np.finfo(float).eps) * max(abs(a), abs(b))


def _unique(v):
    seen, out = set(), []
    for x in v:
        if x not in seen:
            seen.add(x)
            out.append(x)
    return out


def select_one(C, vlist, RSS0, tol=1e-12):
    """group.jl:2070-2080."""
    dC = np.diag(C).copy()
    rs = np.sum(C ** 2, axis=0) / dC
    imax = int(np.argmax(rs))                                      # findmax: first maximum
    vmin = np.sum(dC) - rs[imax]
    residC = C - np.outer(C[:, imax], C[:, imax]) / C[imax, imax]
    index = vlist[imax]
    nzero = np.flatnonzero(np.diag(residC) > tol)
    R2 = 1 - vmin / RSS0
    return index, R2, residC[np.ix_(nzero, nzero)], [vlist[i] for i in nzero]


def select_best_rss_subset(C, k, r2_threshold=1 - 1e-12):
    """group.jl:2081-2098 (returns 0-based indices, most important first)."""
    C = np.array(C, dtype=np.float64)
    p = C.shape[1]
    indices = []
    vlist = list(range(p))
    for _ in range(k):
        idx, r2, C, vlist = select_one(C, vlist, p)
        indices.append(idx)
        if r2 > r2_threshold:
            break
        if len(vlist) == 0:
            break
    return indices


def prioritize_variants(index, priority_vars):
    """group.jl:2050-2056."""
    first = [index.index(v) for v in priority_vars]
    second = [i for i in range(len(index)) if i not in first]
    return [index[i] for i in first] + [index[i] for i in second]


def choose_group_reps(Sigma, groups, threshold=0.5, prioritize_idx=None, Sigma_inv=None):
    """group.jl:1929-2033.  Sigma: correlation matrix (upper triangle trusted).  Returns sorted 0-based indices."""
    Sigma = np.triu(Sigma) + np.triu(Sigma, 1).T
    groups = list(np.asarray(groups))
    p = Sigma.shape[0]
    if not all(_isapprox(x, 1.0) for x in np.diag(Sigma)):
        raise ValueError("Σ must be scaled to a correlation matrix first.")
    if not (0 < threshold < 1):
        raise ValueError(f"threshold should be in (0, 1) but was {threshold}")
    # boundary case: linearly dependent columns (:1938-1975)
    dependent = []
    for i in range(p):
        for j in range(i + 1, p):
            dep = True
            for k in range(i + 1):
                if not _isapprox(Sigma[k, i], Sigma[k, j]):
                    dep = False
                    break
            if dep:
                if prioritize_idx is None:
                    dependent.append(j)
                else:
                    dependent.append(i if j in prioritize_idx else j)
    dependent = _unique(dependent)
    if dependent:
        indep = [i for i in range(p) if i not in dependent]
        sub = Sigma[np.ix_(indep, indep)]
        gsub = [groups[i] for i in indep]
        pr = None
        if prioritize_idx is not None:
            pr = [idx - sum(1 for d in dependent if d < idx) for idx in prioritize_idx]
        reps_sub = choose_group_reps(sub, gsub, threshold=threshold, prioritize_idx=pr)
        return [indep[i] for i in reps_sub]
    # main algorithm (:1977-2032)
    Sinv = np.linalg.inv(Sigma) if Sigma_inv is None else Sigma_inv
    reps = []
    for g in _unique(groups):
        gidx = [i for i in range(p) if groups[i] == g]
        O = [i for i in range(p) if groups[i] != g]
        gs = len(gidx)
        if gs == 1:
            reps.append(gidx[0])
            continue
        index = select_best_rss_subset(Sigma[np.ix_(gidx, gidx)], gs)
        if prioritize_idx is not None:
            ordered = [gidx[i] for i in index]
            pg = [ordered.index(v) for v in prioritize_idx if v in ordered]     # indexin(prioritize_idx, group_idx[index])
            index = prioritize_variants(index, [index[i] for i in pg])
        idxS = [gidx[i] for i in index]
        R = [idxS[0]]
        reps.append(idxS[0])
        for i in range(1, gs):
            if i >= len(idxS):
                break                                              # `indexΣ[i]` would throw in the reference
            RO = R + [o for o in O if o not in R]
            Rc = [v for v in idxS if v not in R]
            ROc = [v for v in range(p) if v not in RO]
            S_RR_inv = np.linalg.inv(Sigma[np.ix_(R, R)])
            L = np.linalg.cholesky(Sinv[np.ix_(ROc, ROc)])
            Xm = np.linalg.solve(L, Sinv[np.ix_(ROc, RO)])
            S_RORO_inv = Sinv[np.ix_(RO, RO)] - Xm.T @ Xm
            ratio = 0.0
            for j in Rc:
                s_Rj = Sigma[R, j]
                s_ROj = Sigma[RO, j]
                ratio += (s_Rj @ S_RR_inv @ s_Rj) / (s_ROj @ S_RORO_inv @ s_ROj)
            ratio /= len(Rc)
            if ratio > threshold:
                break
            R.append(idxS[i])
            reps.append(idxS[i])
    return sorted(reps)


def cond_indep_corr(Sigma, groups, group_reps):
    """group.jl:205-240."""
    p = Sigma.shape[0]
    groups = np.asarray(groups)
    reps = list(group_reps)
    non = [i for i in range(p) if i not in reps]
    b1, b2 = np.zeros((p, p)), np.zeros((p, p))
    for g in _unique(list(groups)):
        gr = [r for r in reps if groups[r] == g]
        gn = [i for i in range(p) if groups[i] == g and i not in gr]
        inv_rr = np.linalg.inv(Sigma[np.ix_(gr, gr)])
        s_rc = Sigma[np.ix_(gr, gn)]
        b1[np.ix_(gr, gn)] = inv_rr @ s_rc
        b2[np.ix_(gn, gn)] = Sigma[np.ix_(gn, gn)] - s_rc.T @ inv_rr @ s_rc
    new = np.zeros((p, p))
    S11 = Sigma[np.ix_(reps, reps)]
    new[np.ix_(reps, reps)] = S11
    d12 = b1[np.ix_(reps, non)]
    new[np.ix_(reps, non)] = S11 @ d12
    new[np.ix_(non, reps)] = new[np.ix_(reps, non)].T
    new[np.ix_(non, non)] = b2[np.ix_(non, non)] + d12.T @ S11 @ d12
    return new


def solve_s_graphical_group(Sigma, groups, group_reps, method, m=1, V=None, **kw):
    """group.jl:268-303.  Returns (S, D, obj)."""
    Sigma = np.triu(Sigma) + np.triu(Sigma, 1).T
    p = Sigma.shape[0]
    groups = np.asarray(groups)
    reps = list(group_reps)
    non = [i for i in range(p) if i not in reps]
    S11 = Sigma[np.ix_(reps, reps)]
    S12 = Sigma[np.ix_(reps, non)]
    S22 = Sigma[np.ix_(non, non)]
    S, _, obj = go.solve_s_group(S11, groups[reps], method, m=m, V=V, **kw)
    inv11 = np.linalg.inv(S11)
    W = inv11 @ S12
    SW = S @ W
    D = np.empty((p, p))
    D[np.ix_(reps, reps)] = S
    D[np.ix_(reps, non)] = SW
    D[np.ix_(non, reps)] = SW.T
    D[np.ix_(non, non)] = S22 - (S12.T @ inv11 @ S12) + (W.T @ S @ W)
    D[np.abs(D) < 1e-10] = 0.0
    return S, D, obj


def modelX_gaussian_rep_group_knockoffs(X, method, groups, mu, Sigma, m=1, rep_threshold=0.5, enforce_cond_indep=False,
                                        Z=None, V=None, group_reps=None, **kw):
    """group.jl:170-203.  Returns dict(Xko, group_reps, S11, S (= D), obj, sigma_used)."""
    if X.shape[1] != len(groups):
        raise ValueError("Dimensions of X and groups doesn't match")
    Ssym = np.triu(Sigma) + np.triu(Sigma, 1).T
    reps = choose_group_reps(Ssym, groups, threshold=rep_threshold) if group_reps is None else list(group_reps)
    sigma = cond_indep_corr(Ssym, groups, reps) if enforce_cond_indep else Ssym
    S, D, obj = solve_s_graphical_group(sigma, groups, reps, method, m=m, V=V, **kw)
    Dsym = np.triu(D) + np.triu(D, 1).T
    Xko = ko.condition(X, np.asarray(mu), Ssym, Dsym, m=m, Z=Z)
    return dict(Xko=Xko, group_reps=reps, S11=S, S=Dsym, obj=obj, sigma_used=sigma)
