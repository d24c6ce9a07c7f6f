"""cholesky(PositiveFactorizations.Positive, ·) of `condition` (modelX.jl:111, :120) -- the opt-in restatement
(csrc/positive.cu, oracle cholesky_positive; PARITY UNPINNED: the package is not vendored with the reference).
Default behaviour stays KB_ERR_NOT_PD; with kb_set_positive_cholesky the :equi solution (κΣ − S exactly singular) and a
deliberately semi-definite / indefinite matrix factor like the oracle's restatement."""
import numpy as np
import pytest

from conftest import toeplitz
from oracle import knockoffs_oracle as ko

pytestmark = pytest.mark.gpu


@pytest.fixture()
def pctx(kb):
    c = kb.Context(0)
    c.set_positive_cholesky(True)
    yield c
    c.close()


@pytest.mark.parametrize("p,kind", [(1, "pd"), (63, "pd"), (64, "semidef"), (65, "indef"), (200, "semidef"), (300, "indef"), (130, "pd")])
def test_positive_factor_matches_restatement(kb, pctx, p, kind):
    """Well-posed inputs only: an unpivoted LDL' of a singular AND indefinite matrix has rounding-level pivots of arbitrary
    sign.  pd: ordinary factor.  indef: full rank with 3 negative eigenvalues (d carries the inertia).  semidef: the last 7
    variables duplicate the first 7 exactly, so their pivots are rounding noise far below tol and are dropped."""
    rng = np.random.default_rng(p)
    if kind == "semidef":
        G = rng.standard_normal((p - 7, p + 5))
        M = np.vstack([G, G[:7]])
        A = M @ M.T / p
    else:
        B = rng.standard_normal((p, p + 5))
        A = B @ B.T / p + 0.3 * np.eye(p)
        if kind == "indef":
            v = rng.standard_normal((p, 3))
            A -= 4.0 * v @ v.T / p
    L, d, mod = kb.hook_potrf_positive(A, ctx=pctx)
    Lw, dw = ko.cholesky_positive(A)
    if kind == "pd":
        assert mod == 0 and np.all(d == 1) and np.max(np.abs(L - np.linalg.cholesky(A))) < 1e-11
        return
    assert np.array_equal(d, dw) and mod == int(np.sum(dw != 1))
    if kind == "indef":
        assert int(np.sum(d == -1)) == 3 and not np.any(d == 0)          # Sylvester: the inertia of A
    else:
        assert int(np.sum(d == 0)) == 7 and np.all(np.diag(L)[d == 0] == 1.0)
    assert np.max(np.abs(L - Lw)) < 1e-8 * max(1.0, np.abs(Lw).max())
    R = (L * d) @ L.T
    keep = d != 0
    assert np.max(np.abs(R[np.ix_(keep, keep)] - A[np.ix_(keep, keep)])) < 1e-9 * max(1.0, np.abs(A).max())


@pytest.mark.parametrize("m", [1, 3])
def test_equi_condition_default_raises_and_positive_matches_oracle(kb, ctx, pctx, m):
    """:equi makes κΣ − S exactly singular, so the conditional covariance is singular: the reference (Positive) returns
    knockoffs; the library raises by default and reproduces the restated Positive factor when opted in."""
    rng = np.random.default_rng(7 + m)
    n, p = 64, 150
    Sigma = toeplitz(p, 0.5)
    X = np.asfortranarray(rng.standard_normal((n, p)) @ np.linalg.cholesky(Sigma).T)
    Z = np.asfortranarray(rng.standard_normal((n, m * p)))
    s = ko.solve_equi(Sigma, m)
    try:
        want = ko.condition(X, np.zeros(p), Sigma, s, m=m, Z=Z)          # plain Cholesky: may or may not survive rounding
        plain_ok = True
    except np.linalg.LinAlgError:
        plain_ok = False
    want_pos = ko.condition(X, np.zeros(p), Sigma, s, m=m, Z=Z, positive=True)
    try:
        got_default = kb.condition(X, np.zeros(p), Sigma, s, m=m, Z=Z, ctx=ctx)
    except kb.PosDefException:
        got_default = None
    got = kb.condition(X, np.zeros(p), Sigma, s, m=m, Z=Z, ctx=pctx)
    assert np.all(np.isfinite(got))
    # in the one singular direction the two arms may round the last pivot differently; everything else agrees
    err = np.abs(got - want_pos)
    assert np.quantile(err, 0.99) < 1e-7, np.quantile(err, 0.99)
    if got_default is not None and plain_ok:
        assert np.quantile(np.abs(got_default - want), 0.99) < 1e-6


def test_equi_one_shot_with_positive(kb, pctx):
    rng = np.random.default_rng(3)
    n, p, m = 200, 100, 2
    Sigma = toeplitz(p, 0.3)
    X = np.asfortranarray(rng.standard_normal((n, p)) @ np.linalg.cholesky(Sigma).T)
    res = kb.modelX_gaussian_knockoffs(X, "equi", np.zeros(p), Sigma, m=m, seed=4, ctx=pctx)
    assert np.allclose(res.s, ko.solve_equi(Sigma, m), rtol=1e-9)
    assert res.Xko.shape == (n, m * p) and np.all(np.isfinite(res.Xko))
    # PD input: the positive fallback never triggers and the factor is the ordinary one
    res2 = kb.modelX_gaussian_knockoffs(X, "maxent", np.zeros(p), Sigma, m=m, seed=4, ctx=pctx)
    assert pctx.positive_modified == 0 and np.all(np.isfinite(res2.Xko))
