"""CPU oracle (numpy, literal loops) for the marginal feature statistics and the knockoff filter.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (only tests/, smoke() and bench.py's CPU legs may import it).

Restates, line by line (paths relative to /root/reference):
  src/threshold.jl:19-31   threshold            src/threshold.jl:55-75   mk_threshold
  src/threshold.jl:96-120  MK_statistics (m>1)  src/threshold.jl:134-142 MK_statistics (m=1)
  src/fit_lasso.jl:201-257 fit_marginal(y, ko)
Pinned against the reference's own known answers: test/runtests.jl:157-166 (threshold) and :200-228
(MK_statistics) -- tests/golden/reference_known_answers.json keys K7_*, K8_*; see tests/test_oracle_golden.py.
Statistics.median / StatsBase.zscore / std are stdlib-level semantics: median of an even count is
middle(a, b) = a/2 + b/2; std is the corrected (n-1) one.
"""
from __future__ import annotations

import numpy as np


def threshold(w, q, method="knockoff_plus", rej_bounds=10000):
    """threshold.jl:19-31."""
    if not (0 <= q <= 1):
        raise ValueError(f"Target FDR should be between 0 and 1 but got {q}")
    if method == "knockoff":
        offset = 0
    elif method == "knockoff_plus":
        offset = 1
    else:
        raise ValueError(f"method should be :knockoff or :knockoff_plus but was {method}.")
    w = np.asarray(w, dtype=np.float64)
    tau = np.inf
    for i, t in enumerate(np.sort(np.abs(w))[::-1], 1):          # t starts from largest |W|
        num = offset + int(np.count_nonzero(w <= -t))
        den = int(np.count_nonzero(w >= t))
        ratio = num / den if den > 0 else (np.inf if num > 0 else np.nan)
        if ratio <= q and t > 0:
            tau = t
        if i > rej_bounds:
            break
    return float(tau)


def mk_threshold(tau, kappa, m, q, method="knockoff_plus", rej_bounds=10000):
    """threshold.jl:55-75."""
    if not (0 <= q <= 1):
        raise ValueError(f"Target FDR should be between 0 and 1 but got {q}")
    if method != "knockoff_plus":
        raise ValueError("Multiple knockoffs needs to use :knockoff_plus filtering method")
    tau = np.asarray(tau, dtype=np.float64)
    kappa = np.asarray(kappa)
    if len(tau) != len(kappa):
        raise ValueError("Length of τ and κ should be the same")
    tau_hat = np.inf
    offset = 1 / m
    for i, t in enumerate(np.sort(tau)[::-1], 1):
        ge = tau >= t
        numer = int(np.count_nonzero(ge & (kappa >= 1)))
        denom = int(np.count_nonzero(ge & (kappa == 0)))
        ratio = (offset + offset * numer) / max(1, denom)
        if ratio <= q and 0 < t < tau_hat:
            tau_hat = t
        if i > rej_bounds:
            break
    return float(tau_hat)


def _median(v):
    """Statistics.median: middle(a, b) = a/2 + b/2 for an even count."""
    s = np.sort(np.asarray(v, dtype=np.float64))
    n = len(s)
    if n % 2:
        return float(s[n // 2])
    return float(s[n // 2 - 1] / 2 + s[n // 2] / 2)


def MK_statistics(T0, Tk):
    """threshold.jl:96-120 when Tk is a list of m > 1 vectors -> (kappa, tau, W);
    threshold.jl:134-142 when Tk is a single vector -> W."""
    T0 = np.asarray(T0, dtype=np.float64)
    if isinstance(Tk, np.ndarray) and Tk.ndim == 1 or (isinstance(Tk, (list, tuple)) and np.ndim(Tk[0]) == 0):
        Tk = np.asarray(Tk, dtype=np.float64)
        if len(T0) != len(Tk):
            raise ValueError("Length of T0 should equal length of Tk")
        return np.abs(T0) - np.abs(Tk)
    Tk = [np.asarray(t, dtype=np.float64) for t in Tk]
    p, m = len(T0), len(Tk)
    if any(len(t) != p for t in Tk):
        raise ValueError("Length of T0 should equal all vectors in Tk")
    kappa = np.zeros(p, dtype=np.int64)
    tau = np.zeros(p)
    W = np.zeros(p)
    for i in range(p):
        storage = np.zeros(m + 1)
        storage[0] = abs(T0[i])
        for k in range(m):
            if abs(Tk[k][i]) > abs(T0[i]):
                kappa[i] = k + 1
            storage[k + 1] = abs(Tk[k][i])
        W[i] = (storage[0] - _median(storage[1:])) * (kappa[i] == 0)
        storage = np.sort(storage)[::-1]
        tau[i] = storage[0] - _median(storage[1:])
    return kappa, tau, W


def zscore_cols(A):
    A = np.asarray(A, dtype=np.float64)
    return (A - A.mean(axis=0)) / A.std(axis=0, ddof=1)


def marginal_T(y, X, Xko, m):
    """fit_lasso.jl:214-222: returns (T0, [T1..Tm])."""
    y = np.asarray(y, dtype=np.float64).ravel()
    n, p = X.shape
    ys = (y - y.mean()) / y.std(ddof=1)
    T0 = (zscore_cols(X).T @ ys) ** 2 / n
    Xs = zscore_cols(Xko)
    Tk = [(Xs[:, k * p:(k + 1) * p].T @ ys) ** 2 / n for k in range(m)]
    return T0, Tk


def fit_marginal(y, X, Xko, m, groups=None, fdrs=(0.01, 0.05, 0.1, 0.25, 0.5), filter_method="knockoff_plus"):
    """fit_lasso.jl:201-257 on an existing knockoff.  Returns dict(T0, Tk, W, kappa, tau, taus, selected)."""
    T0, Tk = marginal_T(y, X, Xko, m)
    if groups is not None:
        groups = np.asarray(groups)
        _, first = np.unique(groups, return_index=True)
        uniq = groups[np.sort(first)]                              # unique(groups): order of first appearance
        T0g = np.array([np.sum(np.abs(T0[groups == g])) for g in uniq])
        Tkg = [np.array([np.sum(np.abs(t[groups == g])) for g in uniq]) for t in Tk]
        T0u, Tku = T0g, Tkg
    else:
        T0u, Tku = T0, Tk
    kappa = tau = None
    if m > 1:
        kappa, tau, W = MK_statistics(T0u, Tku)
    else:
        W = MK_statistics(T0u, Tku[0])
    taus, selected = [], []
    for fdr in fdrs:
        t = mk_threshold(tau, kappa, m, fdr) if m > 1 else threshold(W, fdr, filter_method)
        taus.append(t)
        selected.append(np.flatnonzero(W >= t))
    return dict(T0=T0, Tk=Tk, W=W, kappa=kappa, tau=tau, taus=np.array(taus), selected=selected)
