"""CPU oracle (numpy/scipy) for the model-X Gaussian knockoff hot path of Knockoffs.jl v2.0.2.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``knockoffs.jl_b200``) never does.

Every function is a literal restatement of the cited reference lines
(paths relative to /root/reference).  The reference is pure Julia and Julia is
not installed here, so the oracle is pinned instead against the reference's
published known-answer vectors (docs/src/man/modelX/modelX.md:143-208,280-306 and
test/compare_knockpy.ipynb cells 5-9) -- see tests/test_oracle_golden.py.

Third-party arithmetic that is NOT under /root/reference (parity unpinned, see
DESIGN.md): PositiveFactorizations.cholesky(Positive, .) -- restated as plain
Cholesky (equal for PD input); CovarianceEstimation LinearShrinkage -- restated
from its published algorithm; Optim.Brent -- restated in oracle/group_oracle.py.

Storage convention: the reference keeps the *upper* factor U (A = U'U,
``L.factors`` with uplo='U').  We keep the same U as a C-contiguous numpy
array so row i of U (the Givens row loop) is contiguous.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla


# --------------------------------------------------------------------------
# simulators (src/utilities.jl:357-380, 483-508)
# --------------------------------------------------------------------------
def simulate_AR1(p: int, rho: float) -> np.ndarray:
    """Stationary AR(1) correlation matrix, src/utilities.jl:357-380 with ``rho`` given.

    rhos = log(rho), rhos[1] = 0, cumsum, exp(-|ci - cj|)  ==  rho^|i-j|.
    shift_until_PSD!/cov2cor are no-ops for |rho|<1 (matrix is PD, unit diagonal).
    """
    rhos = np.full(p, np.log(rho))
    rhos[0] = 0.0
    c = np.cumsum(rhos)
    return np.exp(-np.abs(c[:, None] - c[None, :]))


def simulate_block_covariance(groups: np.ndarray, rho: float, gamma: float) -> np.ndarray:
    """src/utilities.jl:483-508 with num_v = 0: 1 on the diagonal, rho within a
    group, gamma*rho between groups."""
    groups = np.asarray(groups)
    same = groups[:, None] == groups[None, :]
    S = np.where(same, rho, gamma * rho)
    np.fill_diagonal(S, 1.0)
    return S


# --------------------------------------------------------------------------
# rank-1 factor updates (src/utilities.jl:516-563, 571-625), uplo = 'U'
# --------------------------------------------------------------------------
def lowrankupdate_turbo(U: np.ndarray, v: np.ndarray) -> None:
    """chol(U'U + v v') in place; destroys v.  src/utilities.jl:516-563."""
    n = v.shape[0]
    nz = np.flatnonzero(v)
    idx_start, idx_end = nz[0], nz[-1]
    early = 0
    for i in range(idx_start, n):
        f, g = U[i, i], v[i]
        # LinearAlgebra.givensAlgorithm(f, g) for f > 0 (dlartg): r = hypot, c = f/r, s = g/r
        if g == 0.0:
            c, s, r = 1.0, 0.0, f
        else:
            r = np.hypot(f, g)
            c, s = f / r, g / r
        if abs(s) < 1e-15 and i >= idx_end:          # :535-540
            early += 1
            if early > 10:
                break
        else:
            early = 0
        U[i, i] = r                                    # :543
        row = U[i, i + 1:]
        vj = v[i + 1:]
        new_row = c * row + s * vj                     # :549-551 (old A_ij on the right)
        v[i + 1:] = -s * row + c * vj
        U[i, i + 1:] = new_row


def lowrankdowndate_turbo(U: np.ndarray, v: np.ndarray) -> None:
    """chol(U'U - v v') in place; destroys v.  src/utilities.jl:571-625.
    Raises np.linalg.LinAlgError where the reference throws PosDefException."""
    n = v.shape[0]
    nz = np.flatnonzero(v)
    idx_start, idx_end = nz[0], nz[-1]
    early = 0
    for i in range(idx_start, n):
        Aii = U[i, i]
        s = v[i] / Aii
        if s * s > 1:                                   # :591-593
            raise np.linalg.LinAlgError(f"PosDefException({i + 1})")
        c = np.sqrt(1 - s * s)
        if abs(s) < 1e-15 and i >= idx_end:            # :597-602
            early += 1
            if early > 10:
                break
        else:
            early = 0
        U[i, i] = c * Aii                               # :605
        vj = v[i + 1:]
        new_row = (U[i, i + 1:] - s * vj) / c           # :611-613 (new A_ij used for v)
        v[i + 1:] = -s * new_row + c * vj
        U[i, i + 1:] = new_row


def _chol_upper(A: np.ndarray) -> np.ndarray:
    """cholesky(Symmetric(A)).factors with uplo='U' (LAPACK dpotrf 'U')."""
    return np.ascontiguousarray(sla.cholesky(A, lower=False, check_finite=False))


def _solve_Ut(U, b):      # ldiv!(x, UpperTriangular(L.factors)', b)
    return sla.solve_triangular(U, b, trans="T", lower=False, check_finite=False)


def _solve_U(U, b):       # ldiv!(x, UpperTriangular(L.factors), b)
    return sla.solve_triangular(U, b, trans="N", lower=False, check_finite=False)


# --------------------------------------------------------------------------
# single-knockoff solvers
# --------------------------------------------------------------------------
def solve_equi(Sigma: np.ndarray, m: float = 1) -> np.ndarray:
    """src/utilities.jl:104-111."""
    lam = np.linalg.eigvalsh(Sigma).min()
    return np.full(Sigma.shape[0], min(1.0, (m + 1) / m * lam))


def solve_quadratic(cn, cd, Sjj, m):
    """src/utilities.jl:187-199."""
    a = -cn - cd ** 2 * m ** 2
    b = 2 * (-cn * Sjj + cd * m ** 2)
    c = -cn * Sjj ** 2 - m ** 2
    if a == 0 and c == 0:
        return 0.0
    with np.errstate(all="ignore"):
        disc = np.sqrt(b * b - 4 * a * c)
        x1 = (-b + disc) / (2 * a)
        x2 = (-b - disc) / (2 * a)
        d = x1 if (-Sjj < x1 < 1.0 / cd) else x2
    if np.isinf(d) or np.isnan(d):
        raise FloatingPointError("delta_j is Inf/NaN")   # :195-196
    return float(d)


def solve_MVR(Sigma, m=1, niter=100, tol=1e-6, lambda_min=1e-6, s_init=None, trace=None):
    """src/utilities.jl:124-175."""
    p = Sigma.shape[0]
    kap = (m + 1) / m
    s = (solve_equi(Sigma, m) if s_init is None else np.array(s_init, dtype=float)).copy()
    U = _chol_upper(kap * Sigma - np.diag(s) + lambda_min * np.eye(p))     # :140
    ej = np.zeros(p)
    sweeps = 0
    for _ in range(niter):
        sweeps += 1
        max_delta = 0.0
        for j in range(p):
            ej[:] = 0
            ej[j] = 1
            vd = _solve_Ut(U, ej)                      # :152 (and first half of :149)
            vn = _solve_U(U, vd)                       # :149
            cn = -np.sum(vn * vn)
            cd = np.sum(vd * vd)
            dj = solve_quadratic(cn, cd, s[j], m)      # :155
            ub = 1 / cd - lambda_min                   # :157-158
            if dj > ub:
                dj = ub
            if abs(dj) < 1e-15:                        # :159
                continue
            s[j] += dj
            ej[j] = np.sqrt(abs(dj))
            if dj > 0:
                lowrankdowndate_turbo(U, ej)
            else:
                lowrankupdate_turbo(U, ej)
            max_delta = max(max_delta, abs(dj))
        if trace is not None:
            trace.append(max_delta)
        if max_delta < tol:                             # :172
            break
    return s, sweeps


def solve_max_entropy(Sigma, m=1, niter=100, tol=1e-6, lambda_min=1e-6, s_init=None, trace=None):
    """src/utilities.jl:221-280 (incl. quirk Q2: s_j takes the *unclipped* value)."""
    p = Sigma.shape[0]
    kap = (m + 1) / m
    s = (solve_equi(Sigma, m) if s_init is None else np.array(s_init, dtype=float)).copy()
    U = _chol_upper(kap * Sigma - np.diag(s) + lambda_min * np.eye(p))     # :237
    x = np.zeros(p)
    sweeps = 0
    for _ in range(niter):
        sweeps += 1
        max_delta = 0.0
        for j in range(p):
            yt = kap * Sigma[:, j]
            yt[j] = 0                                   # :243-246
            xs = _solve_Ut(U, yt)                       # :248
            x_l2 = np.sum(xs * xs)
            zeta = kap * Sigma[j, j] - s[j]
            c = (zeta * x_l2) / (zeta + x_l2)           # :251-252
            sj_new = (kap * Sigma[j, j] - c) / 2        # :254
            x[:] = 0
            x[j] = 1
            w = _solve_Ut(U, x)                         # :258
            ub = 1 / np.sum(w * w) - lambda_min
            d = sj_new - s[j]
            if d > ub:
                d = ub
            if abs(d) < 1e-15:
                continue
            s[j] = sj_new                               # :264  (Q2)
            x[:] = 0
            x[j] = np.sqrt(abs(d))
            if d > 0:
                lowrankdowndate_turbo(U, x)
            else:
                lowrankupdate_turbo(U, x)
            max_delta = max(max_delta, abs(d))
        if trace is not None:
            trace.append(max_delta)
        if max_delta < tol:
            break
    return s, sweeps


def solve_sdp_ccd(Sigma, m=1, lam=0.5, mu=0.8, niter=100, tol=1e-6, trace=None):
    """src/utilities.jl:291-343."""
    if not (0 <= mu <= 1):
        raise ValueError("Decay parameter mu must be in [0, 1]")
    if not lam > 0:
        raise ValueError("Barrier coefficient lambda must be > 0")
    p = Sigma.shape[0]
    kap = (m + 1) / m
    s = np.zeros(p)
    U = _chol_upper(kap * Sigma)                        # :309
    x = np.zeros(p)
    sweeps = 0
    for _ in range(niter):
        sweeps += 1
        for j in range(p):
            yt = kap * Sigma[:, j]
            yt[j] = 0
            xs = _solve_Ut(U, yt)
            x_l2 = np.sum(xs * xs)
            zeta = kap * Sigma[j, j] - s[j]
            c = (zeta * x_l2) / (zeta + x_l2)
            sj_new = min(max(kap * Sigma[j, j] - c - lam, 0.0), 1.0)   # :329
            d = s[j] - sj_new                           # :330 (sign flipped)
            if abs(d) < 1e-15:
                continue
            s[j] = sj_new
            x[:] = 0
            x[j] = np.sqrt(abs(d))
            if d > 0:
                lowrankupdate_turbo(U, x)
            else:
                lowrankdowndate_turbo(U, x)
        if trace is not None:
            trace.append(float(np.sum(s)))
        lam *= mu                                        # :339-340
        if lam < tol:
            break
    return s, sweeps


def solve_s(Sigma, method, m=1, **kw):
    """src/utilities.jl:27-51.  Sigma: full symmetric matrix (upper triangle trusted)."""
    if m < 1:
        raise ValueError(f"m should be 1 or larger but was {m}.")
    Sigma = np.triu(Sigma) + np.triu(Sigma, 1).T        # Symmetric(Sigma), uplo = :U
    sig = np.sqrt(np.diag(Sigma))
    iscor = bool(np.all(np.isclose(sig, 1.0, rtol=np.sqrt(np.finfo(float).eps), atol=0)))
    Scor = Sigma if iscor else Sigma / np.outer(sig, sig)
    method = str(method).lstrip(":")
    if method == "equi":
        s = solve_equi(Scor, m)
    elif method == "mvr":
        s, _ = solve_MVR(Scor, m=m, **kw)
    elif method == "maxent":
        s, _ = solve_max_entropy(Scor, m=m, **kw)
    elif method == "sdp_ccd":
        s, _ = solve_sdp_ccd(Scor, m=m, **kw)
    else:
        raise ValueError(f"Method must be one of mvr/maxent/sdp_ccd/equi but was {method}")
    if not iscor:
        s = s * sig ** 2
    return s


# --------------------------------------------------------------------------
# conditional sampling (src/modelX.jl:96-122), Z injected
# --------------------------------------------------------------------------
def cholesky_positive(A):
    """cholesky(PositiveFactorizations.Positive, Symmetric(A)).L  (src/modelX.jl:111, :120).

    PARITY UNPINNED: PositiveFactorizations.jl is not under /root/reference.  Restated from its published algorithm as
    recalled: unpivoted LDL' with d_j in {+1, -1, 0}; |pivot| > tol: d_j = sign(pivot), L_jj = sqrt|pivot|,
    L_ij = B_ij d_j / L_jj, trailing B -= d_j l l'; otherwise d_j = 0, L_jj = 1, column dropped; tol = 10 n eps max|diag A|;
    the returned factor is L (D replaced by the identity).  For positive definite A this IS the Cholesky factor.
    Returns (L, d)."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    B = np.triu(A) + np.triu(A, 1).T
    L = np.zeros((n, n))
    d = np.zeros(n)
    tol = 10.0 * n * np.finfo(float).eps * np.max(np.abs(np.diag(B)))
    for j in range(n):
        Bjj = B[j, j]
        if abs(Bjj) > tol:
            d[j] = np.sign(Bjj)
            sj = np.sqrt(abs(Bjj))
            L[j, j] = sj
            l = B[j + 1:, j] * (d[j] / sj)
            L[j + 1:, j] = l
            B[j + 1:, j + 1:] -= d[j] * np.outer(l, l)
        else:
            L[j, j] = 1.0
    return L, d


def cond_factor(Sigma, S, m=1, lean=False, positive=False):
    """Returns (SigmaInvS, L) of src/modelX.jl:106-120.  S: vector (Diagonal) or dense matrix.
    L is the *lower* factor of C (m = 1) or of the pm x pm Sigma-tilde (m > 1)."""
    p = Sigma.shape[0]
    Smat = np.diag(S) if np.ndim(S) == 1 else np.asarray(S)
    Sinv = np.linalg.inv(Sigma)                          # :106
    Sinv = (Sinv + Sinv.T) / 2
    SiS = Sinv @ Smat                                    # :107
    C = 2 * Smat - Smat @ SiS                            # :108
    if m == 1:
        Csym = np.triu(C) + np.triu(C, 1).T              # Symmetric(C)
        if positive:
            try:
                return SiS, np.linalg.cholesky(Csym)
            except np.linalg.LinAlgError:
                return SiS, cholesky_positive(Csym)[0]
        return SiS, np.linalg.cholesky(Csym)
    if lean:
        # same matrix as below without the 4 pm x pm temporaries (pm = 25000 at BASELINE configs[2]): Symmetric(St)
        # reads the upper triangle of St, i.e. whole C - S blocks above the block diagonal, their transposes below it,
        # and the upper triangle of (C - S) + S on it.  Factored in place.
        B = C - Smat
        D = B + Smat
        D = np.triu(D) + np.triu(D, 1).T
        St = np.empty((m * p, m * p))
        for k in range(m):
            for l in range(m):
                St[k * p:(k + 1) * p, l * p:(l + 1) * p] = D if k == l else (B if k < l else B.T)
        return SiS, sla.cholesky(St, lower=True, overwrite_a=True, check_finite=False)
    St = np.tile(C - Smat, (m, m))                       # :116
    for k in range(m):
        St[k * p:(k + 1) * p, k * p:(k + 1) * p] += Smat # :117
    St = np.triu(St) + np.triu(St, 1).T
    if positive:
        try:
            return SiS, np.linalg.cholesky(St)
        except np.linalg.LinAlgError:
            return SiS, cholesky_positive(St)[0]
    return SiS, np.linalg.cholesky(St)


def condition(X, mu, Sigma, S, m=1, Z=None, rng=None, lean=False, positive=False):
    """src/modelX.jl:96-122 with Z = randn(n, m*p) injected.  Note quirk Q1:
    the noise is Z @ L with L *lower* (literal reference)."""
    n, p = X.shape
    m = int(m)
    if m < 1:
        raise ValueError(f"m should be 1 or larger but was {m}.")
    if Z is None:
        Z = (rng or np.random.default_rng()).standard_normal((n, m * p))
    SiS, L = cond_factor(Sigma, S, m, lean=lean, positive=positive)
    mui = X - (X - mu[None, :]) @ SiS
    if m == 1:
        return mui + Z @ L                               # :112
    return np.tile(mui, (1, m)) + Z @ L                  # :118-121


def cov_shrink_lw(X):
    """cov(LinearShrinkage(DiagonalUnequalVariance(), :lw), X) as used at src/modelX.jl:43,47.  CovarianceEstimation.jl
    is NOT under /root/reference (parity unpinned): restated from the published Ledoit-Wolf formula with the
    diagonal (unequal variance) target, uncorrected S = Xc'Xc/n:
    lambda = clamp( sum_{i!=j} Var^(s_ij) / sum_{i!=j} s_ij^2, 0, 1 ),  Var^(s_ij) = ((1/n) sum_k (x_ki x_kj)^2 - s_ij^2) / n."""
    n, p = X.shape
    mu = X.mean(0)
    Xc = X - mu
    S = Xc.T @ Xc / n
    pi = (Xc ** 2).T @ (Xc ** 2) / n - S ** 2
    off = ~np.eye(p, dtype=bool)
    lam = float(np.clip(pi[off].sum() / n / np.sum(S[off] ** 2), 0.0, 1.0))
    Sig = (1 - lam) * S + lam * np.diag(np.diag(S))
    return Sig, mu, lam


def modelX_gaussian_knockoffs(X, method, mu, Sigma, m=1, Z=None, **kw):
    """src/modelX.jl:53-66.  Returns (Xko, s)."""
    s = solve_s(Sigma, method, m=m, **kw)
    Ssym = np.triu(Sigma) + np.triu(Sigma, 1).T
    return condition(X, mu, Ssym, s, m=m, Z=Z), s
