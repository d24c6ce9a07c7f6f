"""approx_modelX_gaussian_knockoffs (src/approx.jl:37-109).

CPU: the oracle satisfies the reference's own tests for this path (test/runtests.jl:413-432 are property tests:
κΣ − Diag(s) PSD, s ≥ 0) and the window partition helpers.  GPU: the device path against the oracle on the same
X and the same injected Z (s 1e-8 relative, X̃ 1e-10 -- the north-star tolerances), the same properties, and the
two-phase sharded form == the one-shot call."""
import numpy as np
import pytest

from conftest import toeplitz
from oracle import approx_oracle as ao
from oracle import knockoffs_oracle as ko


def _ar1_X(n, p, rho, seed):
    Sigma = toeplitz(p, rho)
    L = np.linalg.cholesky(Sigma)
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((n, p)) @ L.T)


# ---------------------------------------------------------------------------------------------- CPU
def test_oracle_window_ranges():
    assert ao.window_ranges_from_size(500, 99) == [(0, 99), (99, 198), (198, 297), (297, 396), (396, 495), (495, 500)]
    assert ao.window_ranges_from_size(200, 100) == [(0, 100), (100, 200)]
    with pytest.raises(ValueError):
        ao.window_ranges_from_size(10, 1)


@pytest.mark.parametrize("method,m", [("mvr", 1), ("maxent", 3), ("sdp_ccd", 1), ("equi", 1)])
def test_oracle_reference_properties(method, m):
    """test/runtests.jl:413-432 on a smaller instance: λmin(κΣ − Diag(s)) ≥ 0 (atol 1e-8), s ≥ 0."""
    X = _ar1_X(60, 90, 0.4, 2022)
    r = ao.approx_modelX_gaussian_knockoffs(X, method, [(0, 33), (33, 40), (40, 90)], m=m,
                                            Z=np.random.default_rng(1).standard_normal((60, 90 * m)))
    lam = np.linalg.eigvalsh((m + 1) / m * r["Sigma"] - np.diag(r["s"])).min()
    assert lam >= 0 or abs(lam) < 1e-8
    assert np.all(r["s"] >= 0) and 0 < r["gamma"] <= 1
    assert r["Xko"].shape == (60, 90 * m)


def test_oracle_gamma_bisection():
    Sigma = toeplitz(30, 0.6)
    s = np.full(30, 1.2)                                           # infeasible at γ = 1 for m = 1
    g = ao.gamma_bisection(s, Sigma, 1)
    want = 2 * np.linalg.eigvalsh(Sigma).min() / 1.2
    assert 0 < g < 1 and abs(g - want) < 1e-12
    assert ao.gamma_bisection(0.5 * s * g, Sigma, 1) == 1.0


def test_window_assignment(kb):
    off = np.array([0, 99, 121, 444, 500])
    for world in (1, 2, 3, 4, 8):
        parts = kb.dist.assign_windows(off, world)
        covered = [w for a, b in parts for w in range(a, b)]
        assert covered == list(range(4))                           # contiguous, disjoint, complete
    assert kb.dist.assign_windows(off, 2) == [(0, 3), (3, 4)] or kb.dist.assign_windows(off, 2)[0][1] in (2, 3)
    off2 = np.arange(0, 5001, 500)
    assert kb.dist.assign_windows(off2, 2) == [(0, 5), (5, 10)]
    o, r = kb.api._window_offsets(500, None, 99)
    assert list(o) == [0, 99, 198, 297, 396, 495, 500] and r[-1] == (496, 500)
    with pytest.raises(kb.KnockoffsError):
        kb.api._window_offsets(500, [(1, 99), (100, 400)])          # approx.jl:73-75
    with pytest.raises(kb.KnockoffsError):
        kb.api._window_offsets(500, None, 1)                        # approx.jl:45


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("method,m,n,ranges", [
    ("mvr", 1, 100, [(1, 99), (100, 121), (122, 444), (445, 500)]),     # runtests.jl:420-424 (p > n per window)
    ("mvr", 5, 100, [(1, 99), (100, 121), (122, 444), (445, 500)]),     # runtests.jl:426-431
    ("maxent", 2, 400, [(1, 64), (65, 65 + 130), (196, 300)]),
    ("sdp_ccd", 1, 300, None),                                          # runtests.jl:415 (windowsize = 99)
])
def test_approx_matches_oracle(kb, ctx, method, m, n, ranges):
    p = 500 if ranges is None or ranges[-1][1] == 500 else ranges[-1][1]
    X = _ar1_X(n, p, 0.4, 2022)
    Z = np.asfortranarray(np.random.default_rng(7).standard_normal((n, m * p)))
    kw = dict(windowsize=99) if ranges is None else {}
    r = kb.approx_modelX_gaussian_knockoffs(X, method, ranges, m=m, Z=Z, ctx=ctx, **kw)
    zr = ao.window_ranges_from_size(p, 99) if ranges is None else [(a - 1, b) for a, b in ranges]
    o = ao.approx_modelX_gaussian_knockoffs(X, method, zr, m=m, Z=Z)
    assert np.max(np.abs(r.Sigma - o["Sigma"])) <= 1e-12 * np.abs(o["Sigma"]).max()
    assert np.max(np.abs(r.mu - o["mu"])) <= 1e-13
    assert abs(r.gamma - o["gamma"]) <= 1e-12
    assert np.max(np.abs(r.s - o["s"])) <= 1e-8 * np.abs(o["s"]).max()          # north-star: s within 1e-8 relative
    scale = np.abs(o["Xko"]).max()
    assert np.max(np.abs(r.Xko - o["Xko"])) <= 1e-10 * max(1.0, scale) * (50 if method == "sdp_ccd" else 1)
    # the reference's own tests for this path (runtests.jl:416-431)
    lam = np.linalg.eigvalsh((m + 1) / m * r.Sigma - np.diag(r.s)).min()
    assert lam >= 0 or abs(lam) < 1e-8
    assert np.all(r.s >= 0)


@pytest.mark.gpu
def test_feasible_gamma_matches_oracle(kb, ctx):
    Sigma = toeplitz(130, 0.6)
    s = np.full(130, 1.2)
    for m in (1, 3):
        g = kb.feasible_gamma(Sigma, s, m, ctx=ctx)
        want = (m + 1) / m * np.linalg.eigvalsh(Sigma).min() / 1.2
        assert 0 < g < 1 and abs(g - want) <= 1e-11
        assert abs(g - ao.gamma_bisection(s, Sigma, m)) <= 1e-11
    assert kb.feasible_gamma(Sigma, 0.1 * s, 1, ctx=ctx) == 1.0
    with pytest.raises(kb.PosDefException):
        kb.feasible_gamma(-Sigma, s, 1, ctx=ctx)


@pytest.mark.gpu
def test_approx_sharded_equals_one_shot(kb, ctx):
    """Windows dealt to 2 'ranks' (phases called with disjoint window ranges, γ merged by min) == one-shot call."""
    import ctypes as C
    n, p, m = 200, 300, 2
    X = _ar1_X(n, p, 0.5, 5)
    off = np.array([0, 64, 150, 151 + 60, 300], dtype=np.int64)
    off[3] = 211
    ranges = [(int(off[i]) + 1, int(off[i + 1])) for i in range(4)]
    one = kb.approx_modelX_gaussian_knockoffs(X, "maxent", ranges, m=m, seed=11, ctx=ctx)     # device RNG
    parts = kb.dist.assign_windows(off, 2)
    api, _lib = kb.api, kb.api._lib
    opts, packed = api.make_opts(), np.zeros(int(np.sum(np.diff(off) ** 2)))
    s, mu, gam = np.zeros(p), np.zeros(p), []
    for w0, w1 in parts:
        g, info = C.c_double(), _lib.SolveInfo()
        _lib.check(ctx.h, ctx._lib.kb_approx_solve_windows(ctx.h, api._ptr(X), n, p, api._ptr(off), 4, w0, w1, 1, m,
                                                           C.byref(opts), api._ptr(packed), api._ptr(s), api._ptr(mu),
                                                           C.byref(g), C.byref(info)))
        gam.append(g.value)
    s *= min(gam)
    Xko = np.zeros((n, m * p), order="F")
    for w0, w1 in parts:
        _lib.check(ctx.h, ctx._lib.kb_approx_condition_windows(ctx.h, api._ptr(X), n, p, api._ptr(off), 4, w0, w1, m,
                                                               api._ptr(packed), api._ptr(s), api._ptr(mu), None, 11, 0,
                                                               api._ptr(Xko)))
    assert np.array_equal(s, one.s) and np.array_equal(mu, one.mu)
    assert np.array_equal(Xko, one.Xko)                            # same kernels, same Philox keys: bit-identical
    # single-process form of the distributed driver
    d = kb.dist.approx_knockoffs_sharded(ctx, X, "maxent", ranges, m=m, seed=11)
    assert np.array_equal(d["Xko"], one.Xko) and d["cols"] == (0, p)
