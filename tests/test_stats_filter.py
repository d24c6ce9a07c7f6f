"""Feature statistics + knockoff filter (src/fit_lasso.jl:188-257, src/threshold.jl).

CPU: the oracle against the reference's own known answers (test/runtests.jl:157-166, 200-228 = golden K7/K8).
GPU: the library against the same known answers and against the oracle on seeded inputs (bit-exact for the
integer / comparison outputs kappa and the selections, 1e-10 relative for the floating-point scores)."""
import numpy as np
import pytest

from oracle import stats_oracle as so
from oracle import knockoffs_oracle as ko


# ---------------------------------------------------------------------------------------------- CPU
def test_oracle_threshold_known_answers(golden):
    for c in golden["K7_threshold_runtests_157_166"]["cases"]:
        assert so.threshold(c["w"], c["q"], c["method"]) == c["tau"]


def test_oracle_mk_statistics_known_answers(golden):
    g = golden["K8_mk_statistics_runtests_200_228"]
    w = so.MK_statistics(g["single"]["beta"], g["single"]["betako"])
    assert np.allclose(w, g["single"]["w"], rtol=1.5e-8, atol=1e-15)
    kappa, tau, w = so.MK_statistics(g["multi"]["T0"], g["multi"]["Tk"])
    assert np.allclose(w, g["multi"]["w"], rtol=1.5e-8, atol=0)
    assert list(kappa) == g["multi"]["kappa"]
    assert list(tau) == g["multi"]["tau"]                          # the reference tests these with ==


def test_oracle_mk_threshold_properties():
    rng = np.random.default_rng(5)
    p, m = 400, 5
    T0 = np.abs(rng.standard_normal(p)) + (np.arange(p) < 60) * 3.0
    Tk = [np.abs(rng.standard_normal(p)) for _ in range(m)]
    kappa, tau, W = so.MK_statistics(T0, Tk)
    prev = np.inf
    for q in (0.5, 0.25, 0.1, 0.05):                               # stricter target -> threshold not smaller
        t = so.mk_threshold(tau, kappa, m, q)
        assert t >= prev or prev == np.inf or t == np.inf
        if np.isfinite(t):
            sel = (W >= t)
            num = 1 / m + np.count_nonzero((kappa >= 1) & (tau >= t)) / m
            assert num / max(1, np.count_nonzero((kappa == 0) & (tau >= t))) <= q + 1e-15
            assert np.all(kappa[sel] == 0)
        prev = t
    with pytest.raises(ValueError):
        so.mk_threshold(tau, kappa, m, 1.5)
    with pytest.raises(ValueError):
        so.threshold(W, -0.1)


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_threshold_known_answers(kb, ctx, golden):
    for c in golden["K7_threshold_runtests_157_166"]["cases"]:
        assert kb.threshold(c["w"], c["q"], c["method"], ctx=ctx) == c["tau"]
    with pytest.raises(kb.KnockoffsError):
        kb.threshold([0.1, 0.2], 1.2, ctx=ctx)                     # threshold.jl:21
    with pytest.raises(kb.KnockoffsError):
        kb.threshold([0.1, 0.2], 0.1, "bogus", ctx=ctx)            # threshold.jl:22-23


@pytest.mark.gpu
def test_mk_statistics_known_answers(kb, ctx, golden):
    g = golden["K8_mk_statistics_runtests_200_228"]
    w = kb.MK_statistics(g["single"]["beta"], g["single"]["betako"], ctx=ctx)
    assert np.allclose(w, g["single"]["w"], rtol=1.5e-8, atol=1e-15)
    kappa, tau, w = kb.MK_statistics(g["multi"]["T0"], g["multi"]["Tk"], ctx=ctx)
    assert np.allclose(w, g["multi"]["w"], rtol=1.5e-8, atol=0)
    assert list(kappa) == g["multi"]["kappa"]
    assert list(tau) == g["multi"]["tau"]


@pytest.mark.gpu
@pytest.mark.parametrize("p,m,seed", [(1, 1, 0), (7, 2, 1), (513, 3, 2), (2000, 5, 3), (3000, 4, 4), (12000, 1, 5)])
def test_mk_statistics_and_thresholds_match_oracle(kb, ctx, p, m, seed):
    rng = np.random.default_rng(100 + seed)
    T0 = rng.standard_normal(p) ** 2 + (rng.random(p) < 0.15) * 4.0
    Tk = [rng.standard_normal(p) ** 2 for _ in range(m)]
    # ties and exact zeros, as marginal scores of duplicated / constant columns produce them
    if p > 6:
        Tk[0][3] = T0[3]
        T0[5] = 0.0
        for t in Tk:
            t[5] = 0.0
    if m == 1:
        W = kb.MK_statistics(T0, Tk[0], ctx=ctx)
        assert np.array_equal(W, so.MK_statistics(T0, Tk[0]))
        for q in (0.01, 0.05, 0.1, 0.25, 0.5, 1.0, 0.0):
            for meth in ("knockoff", "knockoff_plus"):
                for rej in (10000, 50):
                    assert kb.threshold(W, q, meth, rej, ctx=ctx) == so.threshold(W, q, meth, rej)
    else:
        kappa, tau, W = kb.MK_statistics(T0, Tk, ctx=ctx)
        k0, t0, w0 = so.MK_statistics(T0, Tk)
        assert np.array_equal(kappa, k0) and np.array_equal(tau, t0) and np.array_equal(W, w0)
        for q in (0.01, 0.05, 0.1, 0.25, 0.5, 1.0):
            for rej in (10000, 50):
                assert kb.mk_threshold(tau, kappa, m, q, rej_bounds=rej, ctx=ctx) == so.mk_threshold(tau, kappa, m, q, rej_bounds=rej)


def _linear_model(kb, n, p, k, seed, genotype=False):
    Sigma = kb.harness.ar1_sigma(p, 0.5)
    X = kb.harness.gaussian_X(n, Sigma, seed)
    rng = np.random.default_rng(seed + 1)
    if genotype:                                                   # 0/1/2 dosages: non-zero means, discrete columns
        X = np.asfortranarray((X > 0.3).astype(float) + (X > -0.5).astype(float))
    beta = np.zeros(p)
    idx = rng.choice(p, k, replace=False)
    beta[idx] = rng.uniform(0.5, 1.0, k) * rng.choice([-1, 1], k)
    y = X @ beta + rng.standard_normal(n)
    return Sigma, X, y, idx


@pytest.mark.gpu
@pytest.mark.parametrize("n,p,m,genotype", [(300, 40, 1, False), (1000, 150, 3, False), (5000, 64, 2, True), (9001, 33, 1, True)])
def test_marginal_stats_match_oracle(kb, ctx, n, p, m, genotype):
    Sigma, X, y, _ = _linear_model(kb, n, p, 8, 11, genotype)
    rng = np.random.default_rng(3)
    Xko = np.asfortranarray(rng.standard_normal((n, m * p)) * 2.0 + 5.0)   # any matrix: scores are per column
    T0, Tk = kb.marginal_stats(y, X, Xko, m, ctx=ctx)
    o0, ok = so.marginal_T(y, X, Xko, m)
    scale = max(o0.max(), max(t.max() for t in ok))
    assert np.max(np.abs(T0 - o0)) <= 1e-10 * scale               # floating point: 1e-10 relative to the largest score
    for a, b in zip(Tk, ok):
        assert np.max(np.abs(a - b)) <= 1e-10 * scale


@pytest.mark.gpu
@pytest.mark.parametrize("m", [1, 4])
def test_fit_marginal_fused_matches_oracle(kb, ctx, m):
    """fit_marginal(y, X, μ, Σ): the fused call (scores accumulated chunk by chunk next to the sampling GEMMs)
    = oracle knockoffs with the same injected Z, then the oracle's fit_marginal."""
    n, p = 3000, 120
    Sigma, X, y, causal = _linear_model(kb, n, p, 12, 21)
    Z = np.asfortranarray(np.random.default_rng(9).standard_normal((n, m * p)))
    mu = np.zeros(p)
    ctx.set_sample_chunk_rows(1024)                                # 3 chunks: exercises the accumulation across chunks
    try:
        r = kb.fit_marginal(y, X, mu, Sigma, method="maxent", m=m, Z=Z, ctx=ctx)
        r2 = kb.fit_marginal(y, X, mu, Sigma, method="maxent", m=m, Z=Z, keep_knockoffs=False, ctx=ctx)
    finally:
        ctx.set_sample_chunk_rows(0)
    want_X, want_s = ko.modelX_gaussian_knockoffs(X, "maxent", mu, Sigma, m=m, Z=Z)
    assert np.max(np.abs(r.ko.Xko - want_X)) < 1e-10
    o = so.fit_marginal(y, X, want_X, m)
    scale = o["T0"].max()
    assert np.max(np.abs(r.T0 - o["T0"])) <= 1e-10 * scale
    assert np.max(np.abs(r.W - o["W"])) <= 1e-9 * scale
    if m > 1:
        assert np.array_equal(r.kappa, o["kappa"])
        assert np.max(np.abs(r.tau - o["tau"])) <= 1e-9 * scale
    for a, b, ta, tb in zip(r.selected, o["selected"], r.taus, o["taus"]):
        assert np.array_equal(a - 1, b)                            # identical selections (1-based vs 0-based)
        assert (ta == tb) or abs(ta - tb) <= 1e-9 * scale
    # X̃ never copied back: same scores, same selections
    assert r2.ko.Xko is None
    assert np.array_equal(r2.W, r.W) and all(np.array_equal(a, b) for a, b in zip(r2.selected, r.selected))
    # the filter finds signal: at q = 0.5 most selections are causal
    sel = set(r.selected[-1] - 1)
    assert len(sel) > 0 and len(sel & set(causal)) >= 0.4 * len(sel)


@pytest.mark.gpu
def test_fit_marginal_on_group_knockoff(kb, ctx):
    """fit_marginal(y, ko) with ko.groups: scores summed per group in order of first appearance (fit_lasso.jl:224-238)."""
    n, p, m = 800, 60, 2
    Sigma, X, y, _ = _linear_model(kb, n, p, 6, 31)
    groups = np.repeat(np.arange(1, p // 3 + 1), 3)[::-1].copy()   # descending ids: first-appearance order matters
    Z = np.asfortranarray(np.random.default_rng(4).standard_normal((n, m * p)))
    g = kb.modelX_gaussian_group_knockoffs(X, "maxent", groups, np.zeros(p), Sigma, m=m, Z=Z, ctx=ctx)
    r = kb.fit_marginal(y, g, ctx=ctx)
    o = so.fit_marginal(y, X, g.Xko, m, groups=groups)
    assert r.W.shape == (p // 3,)
    scale = np.abs(o["W"]).max() + 1e-300
    assert np.max(np.abs(r.W - o["W"])) <= 1e-9 * scale
    assert np.array_equal(r.kappa, o["kappa"])
    for a, b in zip(r.selected, o["selected"]):
        assert np.array_equal(a - 1, b)


@pytest.mark.gpu
def test_marginal_stats_full_size_properties(kb, ctx):
    """BASELINE-size columns (n = 1e5): size-independent properties of the scores -- invariance under column
    shift / positive rescaling (z-scoring), T = n·cor(x, y)²·((n-1)/n)² for a known correlation, and chunk additivity."""
    n, p = 100000, 24
    rng = np.random.default_rng(77)
    y = rng.standard_normal(n)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    X[:, 0] = y                                                    # cor = 1  ->  T = (n-1)²/n
    X[:, 1] = -3.0 * y + 7.0                                       # cor = -1
    Xko = np.asfortranarray(X * rng.uniform(0.5, 2.0, p)[None, :] + rng.uniform(-50, 50, p)[None, :])
    T0, Tk = kb.marginal_stats(y, X, Xko, 1, ctx=ctx)
    assert np.max(np.abs(T0 - Tk[0])) <= 1e-9 * T0.max()
    assert abs(T0[0] - (n - 1) ** 2 / n) <= 1e-9 * n and abs(T0[1] - (n - 1) ** 2 / n) <= 1e-9 * n
    o0, _ = so.marginal_T(y, X, Xko, 1)
    assert np.max(np.abs(T0 - o0)) <= 1e-10 * o0.max()
