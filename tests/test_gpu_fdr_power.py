"""FDR / power parity (BASELINE.json north_star).  Reference pipeline: fit_lasso (fit_lasso.jl:68-158):
lasso on [X X̃] with CV-chosen lambda, W_j = |beta_j| - |beta~_j| (threshold.jl:134-142), knockoff+ threshold
(threshold.jl:19-31).  glmnet is not in this image, so sklearn's LassoCV is used identically for both arms:
arm A = X̃ from the GPU library, arm B = X̃ from the CPU oracle.  With Z injected the arms must select IDENTICAL
sets; with independent noise (device Philox vs numpy) the mean FDP / power must be statistically indistinguishable."""
import numpy as np
import pytest

from conftest import toeplitz
from oracle import c_oracle, knockoffs_oracle as ko

pytestmark = pytest.mark.gpu

QS = [0.1, 0.25, 0.5]


def knockoff_plus_threshold(W, q):
    """threshold.jl:19-31 with the knockoff+ offset."""
    for t in np.sort(np.abs(W[W != 0])):
        if (1 + np.sum(W <= -t)) / max(1, np.sum(W >= t)) <= q:
            return t
    return np.inf


def lasso_W(X, Xko, y):
    from sklearn.linear_model import LassoCV
    p = X.shape[1]
    XX = np.hstack([X, Xko])
    XX = (XX - XX.mean(0)) / XX.std(0)
    fit = LassoCV(cv=3, alphas=15, max_iter=3000, random_state=0).fit(XX, y - y.mean())
    b = fit.coef_
    return np.abs(b[:p]) - np.abs(b[p:])


def fdp_power(W, beta, q):
    sel = np.flatnonzero(W >= knockoff_plus_threshold(W, q))
    true = np.flatnonzero(beta)
    fdp = np.sum(beta[sel] == 0) / max(1, len(sel))
    return fdp, len(np.intersect1d(sel, true)) / len(true), sel


def _problem(seed, n=400, p=120, k=24, amp=1.0):
    rng = np.random.default_rng(seed)
    Sigma = toeplitz(p, 0.4)
    X = np.asfortranarray(rng.standard_normal((n, p)) @ np.linalg.cholesky(Sigma).T)
    beta = np.zeros(p)
    idx = rng.choice(p, k, replace=False)
    beta[idx] = amp * rng.uniform(0.5, 1.0, k) * rng.choice([-1, 1], k)    # runtests.jl:285-300
    y = X @ beta + rng.standard_normal(n)
    return Sigma, X, beta, y, rng


def test_identical_selections_with_injected_Z(kb, ctx):
    for seed in range(3):
        Sigma, X, beta, y, rng = _problem(seed)
        n, p = X.shape
        Z = np.asfortranarray(rng.standard_normal((n, p)))
        gpu = kb.modelX_gaussian_knockoffs(X, "mvr", np.zeros(p), Sigma, Z=Z, ctx=ctx)
        cpu_X, cpu_s = ko.modelX_gaussian_knockoffs(X, "mvr", np.zeros(p), Sigma, Z=Z)
        Wg, Wc = lasso_W(X, gpu.Xko, y), lasso_W(X, cpu_X, y)
        for q in QS:
            assert np.array_equal(fdp_power(Wg, beta, q)[2], fdp_power(Wc, beta, q)[2])


def test_fdr_and_power_statistically_indistinguishable(kb, ctx):
    from scipy import stats
    reps = 16
    res = {"gpu": [], "cpu": []}
    for r in range(reps):
        Sigma, X, beta, y, rng = _problem(100 + r, amp=0.3)
        n, p = X.shape
        gpu = kb.modelX_gaussian_knockoffs(X, "maxent", np.zeros(p), Sigma, seed=9000 + r, ctx=ctx)      # device Philox
        s = ko.solve_s(Sigma, "maxent")
        cpu_X = ko.condition(X, np.zeros(p), Sigma, s, Z=rng.standard_normal((n, p)))                    # numpy noise
        for arm, Xko in (("gpu", gpu.Xko), ("cpu", cpu_X)):
            W = lasso_W(X, Xko, y)
            res[arm].append([v for q in QS for v in fdp_power(W, beta, q)[:2]])
    g, c = np.array(res["gpu"]), np.array(res["cpu"])
    for col in range(g.shape[1]):
        q = QS[col // 2]
        tstat, pval = stats.ttest_ind(g[:, col], c[:, col], equal_var=False)
        assert not (pval < 0.01), f"column {col} differs: gpu {g[:, col].mean():.3f} cpu {c[:, col].mean():.3f} p={pval:.4f}"
        if col % 2 == 0:
            assert g[:, col].mean() <= q + 0.1          # FDR control (mean FDP at target q, finite-sample slack)
    assert g[:, 5].mean() > 0.3                           # the filter has power at q = 0.5


# ---- SURVEY 8(d) at the planned strength: >= 100 replicates, m = 5, MK_statistics / mk_threshold, p = 500 ------------
MK_FDRS = [0.05, 0.1, 0.25, 0.5]


def _mk_problem(seed, Lsig, n=1000, p=500, k=50, amp=0.25):
    """Sparse linear model of runtests.jl:285-300: k nonzero effects of random sign, y = X beta + N(0, 1)."""
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.standard_normal((n, p)) @ Lsig.T)
    beta = np.zeros(p)
    idx = rng.choice(p, k, replace=False)
    beta[idx] = amp * rng.uniform(0.5, 1.0, k) * rng.choice([-1, 1], k)
    y = X @ beta + rng.standard_normal(n)
    return X, beta, y, rng


def _fdp_power_sets(selected_1based, beta):
    sel = np.asarray(selected_1based, dtype=np.int64) - 1
    true = np.flatnonzero(beta)
    return np.sum(beta[sel] == 0) / max(1, len(sel)), len(np.intersect1d(sel, true)) / len(true)


def test_mk_filter_m5_identical_selections_with_injected_Z(kb, ctx):
    """Multiple knockoffs (m = 5) through the FUSED kb_fit_marginal (X-tilde scored on the device, never copied back)
    against the oracle's condition + fit_marginal (threshold.jl:55-120, fit_lasso.jl:201-257): identical kappa, identical
    selected sets for every target FDR; tau / W to rounding."""
    from oracle import stats_oracle as so
    p, m = 500, 5
    Sigma = toeplitz(p, 0.4)
    Lsig = np.linalg.cholesky(Sigma)
    s = ko.solve_s(Sigma, "maxent", m=m)
    SiS, Lf = ko.cond_factor(Sigma, s, m)
    for seed in range(6):
        X, beta, y, rng = _mk_problem(seed, Lsig)
        n = X.shape[0]
        Z = np.asfortranarray(rng.standard_normal((n, m * p)))
        got = kb.fit_marginal(y, X, np.zeros(p), Sigma, method="maxent", m=m, fdrs=MK_FDRS, Z=Z, keep_knockoffs=False, ctx=ctx)
        Xko = np.tile(X - X @ SiS, (1, m)) + Z @ Lf
        want = so.fit_marginal(y, X, Xko, m, fdrs=MK_FDRS)
        assert got.ko.Xko is None
        assert np.array_equal(got.kappa, want["kappa"])
        assert np.allclose(got.tau, want["tau"], rtol=1e-9, atol=1e-13) and np.allclose(got.W, want["W"], rtol=1e-9, atol=1e-13)
        for f in range(len(MK_FDRS)):
            assert np.array_equal(np.asarray(got.selected[f]) - 1, want["selected"][f])


def test_mk_filter_m5_fdr_power_100_replicates(kb, ctx):
    """100 replicates with INDEPENDENT noise in the two arms (device Philox vs numpy): mean FDP and power per target FDR
    are statistically indistinguishable (two-sample t-test, alpha = 0.01) and the multiple-knockoff filter controls FDR."""
    from scipy import stats
    from oracle import stats_oracle as so
    p, m, reps = 500, 5, 100
    Sigma = toeplitz(p, 0.4)
    Lsig = np.linalg.cholesky(Sigma)
    s = ko.solve_s(Sigma, "maxent", m=m)
    SiS, Lf = ko.cond_factor(Sigma, s, m)
    res = {"gpu": [], "cpu": []}
    for r in range(reps):
        X, beta, y, rng = _mk_problem(1000 + r, Lsig)
        n = X.shape[0]
        got = kb.fit_marginal(y, X, np.zeros(p), Sigma, method="maxent", m=m, fdrs=MK_FDRS, seed=77000 + r,
                              keep_knockoffs=False, ctx=ctx)
        Xko = np.tile(X - X @ SiS, (1, m)) + rng.standard_normal((n, m * p)) @ Lf
        want = so.fit_marginal(y, X, Xko, m, fdrs=MK_FDRS)
        res["gpu"].append([v for f in range(len(MK_FDRS)) for v in _fdp_power_sets(got.selected[f], beta)])
        res["cpu"].append([v for f in range(len(MK_FDRS)) for v in _fdp_power_sets(want["selected"][f] + 1, beta)])
    g, c = np.array(res["gpu"]), np.array(res["cpu"])
    print("mean FDP/power per target (gpu):", np.round(g.mean(0), 3), "(cpu):", np.round(c.mean(0), 3))
    for col in range(g.shape[1]):
        q = MK_FDRS[col // 2]
        if g[:, col].std() + c[:, col].std() > 0:
            tstat, pval = stats.ttest_ind(g[:, col], c[:, col], equal_var=False)
            assert not (pval < 0.01), f"column {col}: gpu {g[:, col].mean():.3f} cpu {c[:, col].mean():.3f} p={pval:.4f}"
        if col % 2 == 0:
            # FDR control of the multiple-knockoff filter: mean FDP <= q up to the Monte-Carlo error of 100 replicates
            assert g[:, col].mean() <= q + 3 * g[:, col].std() / np.sqrt(reps) + 0.01
    assert g[:, 7].mean() > 0.3 and g[:, 5].mean() > 0.15      # power at q = 0.5 / 0.25
