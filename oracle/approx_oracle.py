"""CPU oracle (numpy) for approx_modelX_gaussian_knockoffs -- src/approx.jl:37-109.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (only tests/, smoke() and bench.py's CPU legs may import it).

Literal restatement: the covariance is estimated window by window, the windows' s are concatenated, the dense
block-diagonal Σ is ASSEMBLED (approx.jl:93) and `condition` runs on the full p x p matrices exactly as the
reference does (the device path never forms them).  Third-party pieces that are not under /root/reference
(parity unpinned): CovarianceEstimation's LinearShrinkage (see knockoffs_oracle.cov_shrink_lw) and Roots.fzero
(approx.jl:97) -- restated as the plain bisection it performs on f(γ) = 1 − γ if λmin(κΣ − γD) > 0 else −Inf:
γ = 1 when that is feasible, otherwise bisection down to adjacent floating-point numbers, feasible end returned.
The reference's own tests for this path are property tests (test/runtests.jl:413-432: κΣ − Diag(s) is PSD, s ≥ 0),
checked in tests/test_approx.py for both the oracle and the device path.
"""
from __future__ import annotations

import numpy as np

from . import knockoffs_oracle as ko


def window_ranges_from_size(p: int, windowsize: int):
    """approx.jl:46-58 (0-based half-open ranges)."""
    if not windowsize > 1:
        raise ValueError(f"windowsize should be > 1 but was {windowsize}")
    windows = -(-p // windowsize)
    return [(w * windowsize, p if w == windows - 1 else (w + 1) * windowsize) for w in range(windows)]


def feasible(gamma, s, Sigma, m):
    """approx.jl:105-109: λmin((m+1)/m Σ − γ Diag(s)) > 0."""
    return np.linalg.eigvalsh((m + 1) / m * Sigma - np.diag(gamma * s)).min() > 0


def gamma_bisection(s, Sigma, m):
    """γ = fzero(γ -> f(γ, s, Σ, m), 0, 1) (approx.jl:97)."""
    if feasible(1.0, s, Sigma, m):
        return 1.0
    lo, hi = 0.0, 1.0
    while True:
        mid = 0.5 * (lo + hi)
        if not (lo < mid < hi):
            return lo
        if feasible(mid, s, Sigma, m):
            lo = mid
        else:
            hi = mid


def approx_modelX_gaussian_knockoffs(X, method, window_ranges, m=1, Z=None, **kw):
    """approx.jl:62-102.  Returns dict(Xko, s, Sigma, mu, gamma, blocks)."""
    n, p = X.shape
    if sum(b - a for a, b in window_ranges) != p:
        raise ValueError("window_ranges have wrong total dimension")
    blocks, ss = [], []
    for a, b in window_ranges:
        Sig_w, _, _ = ko.cov_shrink_lw(X[:, a:b])                  # :84
        blocks.append(Sig_w)
        ss.append(ko.solve_s(Sig_w, method, m=m, **kw))            # :86
    Sigma = np.zeros((p, p))
    for (a, b), B in zip(window_ranges, blocks):                   # :93
        Sigma[a:b, a:b] = B
    mu = X.mean(axis=0)                                            # :94
    s = np.concatenate(ss)                                         # :95
    gamma = min(gamma_bisection(s[a:b], B, m) for (a, b), B in zip(window_ranges, blocks))   # :97 on a block-diagonal Σ
    s = s * gamma                                                  # :98
    Xko = ko.condition(X, mu, Sigma, s, m=m, Z=Z)                  # :100
    return dict(Xko=Xko, s=s, Sigma=Sigma, mu=mu, gamma=gamma, blocks=blocks)
