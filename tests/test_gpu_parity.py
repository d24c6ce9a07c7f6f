"""GPU parity tests (run with -m gpu on the B200 box).  Every call goes through the C ABI
(libknockoffs_b200.so via ctypes) and is compared with the CPU oracle on the same seeded inputs.
Tolerances (BASELINE.json north_star): s 1e-8 relative, factors 1e-8, X̃ 1e-10, objectives 1e-8."""
import numpy as np
import pytest

from conftest import toeplitz
from oracle import c_oracle, knockoffs_oracle as ko

pytestmark = pytest.mark.gpu

S_RTOL = 1e-8
XKO_TOL = 1e-10


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


# ---- dense building blocks -------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,ta,tb", [(128, 128, 16, 0, 0), (257, 131, 77, 0, 0), (300, 200, 129, 1, 0),
                                         (64, 500, 333, 0, 1), (511, 129, 260, 1, 1), (1, 1, 1, 0, 0)])
def test_dgemm(kb, ctx, M, N, K, ta, tb):
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((K, M) if ta else (M, K))
    B = rng.standard_normal((N, K) if tb else (K, N))
    C0 = rng.standard_normal((M, N))
    got = kb.hook_dgemm(A, B, C0, transa=ta, transb=tb, alpha=0.7, beta=-1.3, ctx=ctx)
    want = 0.7 * (A.T if ta else A) @ (B.T if tb else B) - 1.3 * C0
    assert np.max(np.abs(got - want)) < 1e-11 * max(1, K)


@pytest.mark.parametrize("p", [1, 5, 127, 128, 129, 500, 777])
def test_potrf_and_inverse(kb, ctx, p):
    rng = np.random.default_rng(p)
    B = rng.standard_normal((p, p + 3))
    A = B @ B.T / p + 0.5 * np.eye(p)
    Au = np.triu(A) + np.tril(rng.standard_normal((p, p)), -1)      # garbage below: only the upper triangle is read
    L = kb.hook_potrf(Au, ctx=ctx)
    assert np.max(np.abs(L - np.linalg.cholesky(A))) < 1e-11
    Ai = kb.hook_spd_inverse(Au, ctx=ctx)
    assert np.max(np.abs(Ai - np.linalg.inv(A))) < 1e-9
    assert np.max(np.abs(Ai - Ai.T)) == 0.0


def test_potrf_not_pd_raises(kb, ctx):
    A = np.eye(200); A[150, 150] = -1.0
    with pytest.raises(kb.PosDefException):
        kb.hook_potrf(A, ctx=ctx)


def test_eigmin(kb, ctx):
    for p, rho in [(50, 0.4), (300, 0.5), (1000, 0.4)]:
        S = toeplitz(p, rho)
        lam = kb.eigmin(S, ctx=ctx)
        assert abs(lam - np.linalg.eigvalsh(S).min()) < 1e-11


# ---- CCD solvers --------------------------------------------------------------------------------------
def _sigma(kind, p):
    if kind == "ar1":
        return toeplitz(p, 0.5)
    g = np.repeat(np.arange((p + 9) // 10), 10)[:p]
    return np.asfortranarray(ko.simulate_block_covariance(g, 0.5, 0.25))


@pytest.mark.parametrize("kind,p,method,m", [
    ("ar1", 500, "mvr", 1),          # BASELINE config 1
    ("ar1", 500, "maxent", 1), ("ar1", 300, "sdp_ccd", 1),
    ("block", 400, "mvr", 5), ("block", 400, "maxent", 5), ("block", 250, "sdp_ccd", 5),
    ("ar1", 63, "mvr", 1), ("ar1", 64, "maxent", 2), ("ar1", 65, "sdp_ccd", 1), ("ar1", 1, "maxent", 1),
    ("ar1", 2, "maxent", 1), ("block", 130, "mvr", 3),
])
@pytest.mark.parametrize("mode,refresh", [("block", -1), ("block", 1), ("block", 0), ("stream", 1), ("stream", 0)])
def test_solve_s_matches_oracle(kb, ctx, kind, p, method, m, mode, refresh):
    """Both kernels (blocked rank-64 delayed updates = AUTO, rank-1 streaming) against the C oracle."""
    Sigma = _sigma(kind, p)
    s0 = ko.solve_equi(Sigma, m)
    want = c_oracle.solve_ccd(Sigma, method, m=m, s_init=s0)
    got, info = kb.solve_s(Sigma, method, m=m, s_init=s0, ctx=ctx, return_info=True, refresh_every=refresh, mode=mode)
    assert info["mode"] == mode
    assert info["sweeps"] == want["sweeps"]
    assert rel(got, want["s"]) < S_RTOL
    if method != "sdp_ccd":
        # per-sweep max|δ|: the equi starting point makes the first factorisation nearly singular (λmin(A) = 1e-6), so
        # implementations that differ in rounding leave sweep 1 ~1e-9 apart in s and re-converge (an every-sweep
        # float64 numpy restatement shows the same 1e-5 relative spread on the late, tiny δ's)
        assert np.allclose(info["trace"], want["trace"], rtol=1e-4, atol=1e-8)
    assert info["updates"] == want["n_updates"]


def test_solve_s_default_init_and_covariance_scaling(kb, ctx):
    """NULL s_init -> solve_equi on the device; covariance input -> cov2cor and s .*= sigma^2 (utilities.jl:31-49)."""
    p = 300
    Sigma = toeplitz(p, 0.5)
    d = np.linspace(0.5, 2.0, p)
    Cov = np.asfortranarray(Sigma * np.outer(d, d))
    want = ko.solve_s(Cov, "maxent", m=1)
    got = kb.solve_s(Cov, "maxent", m=1, ctx=ctx)
    assert rel(got, want) < S_RTOL
    # only the upper triangle is trusted, inputs are never written
    junk = np.asfortranarray(np.triu(Cov) + np.tril(np.full((p, p), 7.0), -1))
    keep = junk.copy()
    got2 = kb.solve_s(junk, "maxent", m=1, ctx=ctx)
    assert np.array_equal(junk, keep) and np.array_equal(got, got2)


def test_golden_K1_on_gpu(kb, ctx, golden):
    """docs/src/man/modelX/modelX.md:143-169: 16-digit MVR vector, Toeplitz rho=0.4, p=1000, all defaults."""
    g = golden["K1_mvr_ar1_rho0.4_p1000_m1"]
    s = kb.solve_s(toeplitz(1000, 0.4), "mvr", ctx=ctx)
    h, t = np.array(g["head"]), np.array(g["tail"])
    assert np.max(np.abs(s[:len(h)] - h)) < 1e-9 and np.max(np.abs(s[-len(t):] - t)) < 1e-9


def test_reference_property_tests(kb, ctx):
    """runtests.jl:773-808 ('multiple knockoffs'): PSD of (m+1)/m Sigma - diag(s), 0 <= s <= 1."""
    Sigma = toeplitz(500, 0.4)
    for method, m in [("mvr", 3), ("maxent", 5), ("sdp_ccd", 5)]:
        s = kb.solve_s(Sigma, method, m=m, ctx=ctx)
        assert np.all(s >= 0) and np.all(s <= 1 + 1e-12)
        assert np.linalg.eigvalsh((m + 1) / m * Sigma - np.diag(s)).min() > -1e-8
    # runtests.jl:349-365 keyword pass-through
    a = kb.solve_s(Sigma, "sdp_ccd", ctx=ctx, λ=0.7, μ=0.7)
    b = kb.solve_s(Sigma, "sdp_ccd", ctx=ctx, λ=0.9, μ=0.9)
    assert np.max(np.abs(a - b)) < 0.05


def test_error_behaviour(kb, ctx):
    Sigma = toeplitz(50, 0.4)
    with pytest.raises(kb.KnockoffsError):      # utilities.jl:28
        kb.solve_s(Sigma, "mvr", m=0.5, ctx=ctx)
    with pytest.raises(kb.KnockoffsError):      # utilities.jl:46
        kb.solve_s(Sigma, "nonsense", ctx=ctx)
    with pytest.raises(kb.KnockoffsError):      # utilities.jl:301
        kb.solve_s(Sigma, "sdp_ccd", ctx=ctx, μ=1.5)
    with pytest.raises(kb.KnockoffsError):      # utilities.jl:302
        kb.solve_s(Sigma, "sdp_ccd", ctx=ctx, λ=0.0)
    bad = -np.eye(50)
    with pytest.raises(kb.KnockoffsError):
        kb.solve_s(bad, "mvr", ctx=ctx, s_init=np.full(50, 0.5))
    notpd = np.asfortranarray(np.full((50, 50), 0.999) + 0.001 * np.eye(50))
    with pytest.raises(kb.PosDefException):     # cholesky at utilities.jl:140 throws
        kb.solve_s(notpd, "mvr", ctx=ctx, s_init=np.full(50, 1.0))
    with pytest.raises(kb.KnockoffsError):      # modelX.jl:105
        kb.condition(np.zeros((4, 50)), np.zeros(50), Sigma, np.full(50, 0.3), m=0, ctx=ctx)


def test_context_is_reusable_after_errors(kb):
    """An error return must leave no work in flight and no workspace checked out: the same context then reproduces a
    fresh context's results bit for bit."""
    rng = np.random.default_rng(11)
    p, n, m = 120, 300, 2
    Sigma = toeplitz(p, 0.5)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    want_s = kb.solve_s(Sigma, "maxent", m=m, ctx=kb.Context(0))
    want_X = kb.condition(X, np.zeros(p), Sigma, want_s, m=m, seed=5, ctx=kb.Context(0))
    c = kb.Context(0)
    notpd = np.asfortranarray(np.full((p, p), 0.999) + 0.001 * np.eye(p))
    for _ in range(2):
        with pytest.raises(kb.PosDefException):                       # CCD: first Cholesky fails
            kb.solve_s(notpd, "mvr", ctx=c, s_init=np.full(p, 1.0))
        with pytest.raises(kb.PosDefException):                       # condition: 2S - S inv(Sigma) S indefinite for s = 2 diag
            kb.condition(X, np.zeros(p), Sigma, np.full(p, 2.0), m=m, seed=5, ctx=c)
        with pytest.raises(kb.KnockoffsError):
            kb.modelX_gaussian_knockoffs(X, "nonsense", np.zeros(p), Sigma, m=m, ctx=c)
        with pytest.raises(kb.PosDefException):                       # one-shot entry, error in the middle of the pipeline
            kb.modelX_gaussian_knockoffs(X, "mvr", np.zeros(p), notpd, m=m, ctx=c, s_init=np.full(p, 1.0))
        got_s = kb.solve_s(Sigma, "maxent", m=m, ctx=c)
        got_X = kb.condition(X, np.zeros(p), Sigma, got_s, m=m, seed=5, ctx=c)
        assert np.array_equal(got_s, want_s) and np.array_equal(got_X, want_X)


def test_shapes_are_validated_before_the_c_call(kb, ctx):
    """DimensionMismatch-style errors instead of an out-of-bounds host read behind the raw pointers."""
    p, n = 40, 8
    Sigma = toeplitz(p, 0.4)
    X = np.zeros((n, p))
    s = np.full(p, 0.3)
    SiS, Lb = kb.cond_factor(Sigma, s, m=2, ctx=ctx)
    bad_calls = [
        lambda: kb.solve_s(Sigma, "mvr", s_init=np.full(p - 1, 0.5), ctx=ctx),
        lambda: kb.solve_s(Sigma[:, :-1], "mvr", ctx=ctx),
        lambda: kb.eigmin(Sigma[:-1], ctx=ctx),
        lambda: kb.cond_factor(Sigma, s[:-1], m=1, ctx=ctx),
        lambda: kb.cond_factor(Sigma, np.eye(p - 1), m=1, ctx=ctx),
        lambda: kb.sample(X, np.zeros(p), SiS, Lb, m=3, ctx=ctx),              # 3 blocks given, 5 needed
        lambda: kb.sample(X, np.zeros(p - 1), SiS, Lb, m=2, ctx=ctx),
        lambda: kb.sample(X, np.zeros(p), SiS[:-1, :-1], Lb, m=2, ctx=ctx),
        lambda: kb.modelX_gaussian_knockoffs(X, "mvr", np.zeros(p), Sigma, s_init=np.ones(p + 1), ctx=ctx),
        lambda: kb.modelX_gaussian_group_knockoffs(X, "maxent", np.arange(p) // 2 + 1, np.zeros(p - 1), Sigma, ctx=ctx),
        lambda: kb.modelX_gaussian_group_knockoffs(X, "maxent", np.arange(p) // 2 + 1, np.zeros(p), Sigma[:-1, :-1], ctx=ctx),
        lambda: kb.solve_s_group(Sigma, np.arange(p) // 2 + 1, "maxent", V=np.zeros((p - 1, 3)), ctx=ctx),
        lambda: kb.feasible_gamma(Sigma, s[:-1], ctx=ctx),
        lambda: kb.choose_group_reps(Sigma, np.ones(p - 1, dtype=int), ctx=ctx),
        lambda: kb.cond_indep_corr(Sigma, np.ones(p, dtype=int), [0], ctx=ctx),
        lambda: kb.solve_s_graphical_group(Sigma, np.ones(p, dtype=int), [p + 1], "maxent", ctx=ctx),
    ]
    for i, call in enumerate(bad_calls):
        with pytest.raises(kb.KnockoffsError):
            call()
            pytest.fail(f"bad call {i} was accepted")


# ---- conditional factor + sampling ---------------------------------------------------------------------
@pytest.mark.parametrize("p,m,method", [(200, 1, "maxent"), (130, 3, "mvr"), (257, 5, "sdp_ccd"), (192, 4, "equi09")])
def test_cond_factor_block_columns_equal_the_recursion(kb, ctx, p, m, method):
    """kb_cond_factor_block_dev: block column k from the closed form A_k = ((k+1)/k) S - (1/k) S (k Sigma - (k-1) S)^-1 S
    against the recursion of modelX.jl:116-120 (kb_cond_factor) and against the oracle's dense pm x pm Cholesky factor."""
    import torch
    Sigma = toeplitz(p, 0.5)
    if method == "equi09":
        s = 0.9 * kb.solve_s(Sigma, "equi", m=m, ctx=ctx)
    else:
        s = kb.solve_s(Sigma, method, m=m, ctx=ctx)
    SiS_w, Lb_w = kb.cond_factor(Sigma, s, m=m, ctx=ctx)
    _, L_oracle = ko.cond_factor(Sigma, s, m)
    d = torch.device("cuda", 0)
    dSig = torch.from_numpy(np.ascontiguousarray(Sigma)).to(d)
    ds = torch.from_numpy(s).to(d)
    dSiS = torch.zeros((p, p), dtype=torch.float64, device=d)
    dLb = torch.zeros((2 * m - 1, p, p), dtype=torch.float64, device=d)
    torch.cuda.synchronize()
    for k in range(m, 0, -1):                              # any order: the columns are independent
        kb.dev.cond_factor_block_dev(ctx, dSig.data_ptr(), p, ds.data_ptr(), k, m, dSiS.data_ptr() if k == 1 else 0,
                                     dLb[2 * (k - 1)].data_ptr(), dLb[2 * (k - 1) + 1].data_ptr() if k < m else 0)
    got = np.transpose(dLb.cpu().numpy(), (2, 1, 0))       # device blocks are column-major p x p
    tol = 1e-6 if method == "sdp_ccd" else 1e-9            # :sdp_ccd ends on the feasibility boundary (cond ~ 1e9)
    scale = np.max(np.abs(Lb_w))
    assert np.max(np.abs(got - Lb_w)) / scale < tol
    assert np.max(np.abs(dSiS.cpu().numpy().T - SiS_w)) < 1e-11
    L = kb.dense_factor_from_blocks(np.asfortranarray(got))
    assert np.max(np.abs(L - L_oracle)) / np.max(np.abs(L_oracle)) < (1e-6 if method == "sdp_ccd" else 1e-8)
    with pytest.raises(kb.KnockoffsError):
        kb.dev.cond_factor_block_dev(ctx, dSig.data_ptr(), p, ds.data_ptr(), m + 1, m, 0, dLb[0].data_ptr(), 0)


@pytest.mark.parametrize("p,m", [(200, 1), (130, 3), (257, 2)])
def test_cond_factor_matches_oracle(kb, ctx, p, m):
    Sigma = toeplitz(p, 0.5)
    s = c_oracle.solve_ccd(Sigma, "maxent", m=m, s_init=ko.solve_equi(Sigma, m))["s"]
    SiS_w, L_w = ko.cond_factor(Sigma, s, m)
    SiS, Lb = kb.cond_factor(Sigma, s, m=m, ctx=ctx)
    assert np.max(np.abs(SiS - SiS_w)) < 1e-11
    L = kb.dense_factor_from_blocks(Lb)
    assert np.max(np.abs(L - L_w)) / np.max(np.abs(L_w)) < 1e-8


@pytest.mark.parametrize("n,p,m", [(1000, 500, 1), (333, 130, 3), (4097, 200, 2), (1, 64, 1), (9000, 96, 5),
                                   (257, 131, 3), (65, 33, 4), (4099, 67, 2)])   # odd p / n: unvectorised operand paths
def test_condition_matches_oracle_with_injected_Z(kb, ctx, n, p, m):
    rng = np.random.default_rng(n + p + m)
    Sigma = toeplitz(p, 0.5)
    mu = rng.standard_normal(p)
    X = np.asfortranarray(rng.standard_normal((n, p)) @ np.linalg.cholesky(Sigma).T + mu)
    Z = np.asfortranarray(rng.standard_normal((n, m * p)))
    s = c_oracle.solve_ccd(Sigma, "mvr", m=m, s_init=ko.solve_equi(Sigma, m))["s"]
    want = ko.condition(X, mu, Sigma, s, m=m, Z=Z)
    got = kb.condition(X, mu, Sigma, s, m=m, Z=Z, ctx=ctx)
    assert got.shape == (n, m * p)
    assert np.max(np.abs(got - want)) < XKO_TOL


def test_condition_dense_S(kb, ctx):
    """group path: S dense block-diagonal (group.jl:125 -> modelX.jl:107-108)."""
    rng = np.random.default_rng(5)
    p, n, m = 120, 400, 2
    g = np.repeat(np.arange(p // 5), 5)
    Sigma = np.asfortranarray(ko.simulate_block_covariance(g, 0.6, 0.3))
    S = np.where(g[:, None] == g[None, :], Sigma, 0.0) * 0.4
    X = np.asfortranarray(rng.standard_normal((n, p)))
    Z = np.asfortranarray(rng.standard_normal((n, m * p)))
    want = ko.condition(X, np.zeros(p), Sigma, S, m=m, Z=Z)
    got = kb.condition(X, np.zeros(p), Sigma, S, m=m, Z=Z, ctx=ctx)
    assert np.max(np.abs(got - want)) < XKO_TOL


def test_factor_structure_is_detected(kb, ctx):
    """The sampling stage's short form (two triangular products per copy) rests on M_k = L_kk + N_k with N_k upper
    triangular (Diagonal(s)) or banded below the diagonal (contiguous group blocks); anything else takes the general form."""
    p, m = 120, 3
    g = np.repeat(np.arange(p // 5), 5)
    Sigma = np.asfortranarray(ko.simulate_block_covariance(g, 0.6, 0.3))
    _, Lb = kb.cond_factor(Sigma, np.full(p, 0.3), m=m, ctx=ctx)
    assert kb.hook_factor_band(Lb, ctx=ctx) == 0
    S = np.where(g[:, None] == g[None, :], Sigma, 0.0) * 0.4
    _, Lg = kb.cond_factor(Sigma, S, m=m, ctx=ctx)
    assert kb.hook_factor_band(Lg, ctx=ctx) == 16                  # g − 1 = 4 rows below the diagonal, rounded up to 16
    perm = np.random.default_rng(0).permutation(p)                 # scattered groups: no band
    _, Lp = kb.cond_factor(np.asfortranarray(Sigma[np.ix_(perm, perm)]), np.asfortranarray(S[np.ix_(perm, perm)]), m=m, ctx=ctx)
    assert kb.hook_factor_band(Lp, ctx=ctx) == -1
    want = ko.cond_factor(Sigma, S, m)[1]
    assert np.max(np.abs(kb.dense_factor_from_blocks(Lg) - want)) / np.max(np.abs(want)) < 1e-8


def test_modelX_one_shot_config1(kb, ctx):
    """BASELINE config 1 end to end through kb_modelx_gaussian_knockoffs with Z injected."""
    p, n = 500, 1000
    Sigma = kb.harness.ar1_sigma(p, 0.5)
    X = kb.harness.gaussian_X(n, Sigma, 20260001)
    Z = np.asfortranarray(np.random.default_rng(7).standard_normal((n, p)))
    mu = np.zeros(p)
    want_X, want_s = ko.modelX_gaussian_knockoffs(X, "mvr", mu, Sigma, m=1, Z=Z)
    ko_res = kb.modelX_gaussian_knockoffs(X, "mvr", mu, Sigma, m=1, Z=Z, ctx=ctx)
    assert rel(ko_res.s, want_s) < S_RTOL
    assert np.max(np.abs(ko_res.Xko - want_X)) < XKO_TOL
    assert ko_res.method == "mvr" and ko_res.m == 1 and ko_res.Xko.shape == (n, p)


# ---- device RNG ------------------------------------------------------------------------------------------
def _philox_ref(seed, rows, colpairs):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    r, c = np.meshgrid(rows.astype(np.uint64), colpairs.astype(np.uint64), indexing="ij")
    c0 = (r & 0xFFFFFFFF).astype(np.uint64); c1 = (r >> 32).astype(np.uint64)
    c2 = c.copy(); c3 = np.zeros_like(c)
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)
    for _ in range(10):
        p0 = M0 * c0; p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & 0xFFFFFFFF
        hi1, lo1 = p1 >> 32, p1 & 0xFFFFFFFF
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF; k1 = (k1 + W1) & 0xFFFFFFFF
    a = (c0 << 32) | c1; b = (c2 << 32) | c3
    u1 = ((a >> 11) + 1).astype(np.float64) * 2.0 ** -53
    u2 = (b >> 11).astype(np.float64) * 2.0 ** -53
    rad = np.sqrt(-2.0 * np.log(u1))
    return rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)


def test_device_randn_is_philox_boxmuller_and_shardable(kb, ctx):
    n, nc = 257, 11
    Z = kb.hook_randn(12345678901, 1000, n, nc, ctx=ctx)
    z0, z1 = _philox_ref(12345678901, 1000 + np.arange(n), np.arange((nc + 1) // 2))
    want = np.empty((n, 2 * ((nc + 1) // 2)))
    want[:, 0::2], want[:, 1::2] = z0, z1
    assert np.max(np.abs(Z - want[:, :nc])) < 1e-12
    # a row shard regenerates the same numbers
    Zs = kb.hook_randn(12345678901, 1100, 50, nc, ctx=ctx)
    assert np.array_equal(Zs, Z[100:150])
    big = kb.hook_randn(7, 0, 20000, 100, ctx=ctx).ravel()
    assert abs(big.mean()) < 4 / np.sqrt(big.size) and abs(big.var() - 1) < 0.01
    assert abs(np.mean(big ** 4) - 3) < 0.05


def test_device_sampling_moments(kb, ctx):
    """Z drawn on the device: Cov of the noise part is L'L (reference-literal Z*L)."""
    p, n = 40, 200000
    Sigma = toeplitz(p, 0.5)
    s = c_oracle.solve_ccd(Sigma, "maxent", m=1, s_init=ko.solve_equi(Sigma, 1))["s"]
    SiS, Lb = kb.cond_factor(Sigma, s, ctx=ctx)
    X = np.zeros((n, p), order="F")
    Xko = kb.sample(X, np.zeros(p), SiS, Lb, m=1, Z=None, seed=99, ctx=ctx)
    L = Lb[:, :, 0]
    C = Xko.T @ Xko / n
    assert np.max(np.abs(C - L.T @ L)) < 0.02


# ---- Gram --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,p", [(1000, 500), (70001, 130), (5, 3)])
def test_gram(kb, ctx, n, p):
    rng = np.random.default_rng(n)
    X = np.asfortranarray(rng.standard_normal((n, p)) + rng.standard_normal(p))
    G, mu = kb.gram(X, center=True, ctx=ctx)
    Xc = X - X.mean(0)
    assert np.max(np.abs(mu - X.mean(0))) < 1e-12
    assert np.max(np.abs(G - Xc.T @ Xc / n)) < 1e-10
    G2, _ = kb.gram(X, center=False, denom=n - 1, ctx=ctx)
    assert np.max(np.abs(G2 - X.T @ X / (n - 1))) < 1e-10


# ---- multi-GPU building blocks on one device (the N>1 NCCL path is exercised by bench.py --gpus N) ----------
def test_sharded_gram_dev_equals_single(kb, ctx):
    import torch
    rng = np.random.default_rng(11)
    n, p = 3001, 96
    X = rng.standard_normal((n, p)) + 1.5
    G_acc = torch.zeros((p, p), dtype=torch.float64, device="cuda")
    cs_acc = torch.zeros(p, dtype=torch.float64, device="cuda")
    for r in range(2):
        lo, hi = kb.dist.shard_rows(n, 2, r)
        Xt = torch.from_numpy(np.ascontiguousarray(X[lo:hi].T)).cuda()        # (p, n_local): column-major rows
        G = torch.empty((p, p), dtype=torch.float64, device="cuda")
        cs = torch.empty(p, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        kb.dev.gram_partial_dev(ctx, Xt.data_ptr(), hi - lo, hi - lo, p, G.data_ptr(), cs.data_ptr())
        G_acc += G; cs_acc += cs                                            # what the NCCL all-reduce does
    torch.cuda.synchronize()
    kb.dev.gram_finish_dev(ctx, G_acc.data_ptr(), cs_acc.data_ptr(), n, p, True, 0.0)
    Xc = X - X.mean(0)
    assert np.max(np.abs(G_acc.cpu().numpy() - Xc.T @ Xc / n)) < 1e-10
    assert np.max(np.abs(cs_acc.cpu().numpy() - X.mean(0))) < 1e-12


def test_row_sharded_sampling_equals_single(kb, ctx):
    """Device RNG keyed by the GLOBAL row: two shards with row offsets reproduce the unsharded X~ bit for bit."""
    rng = np.random.default_rng(12)
    n, p, m = 5000, 64, 3
    Sigma = toeplitz(p, 0.5)
    s = c_oracle.solve_ccd(Sigma, "maxent", m=m, s_init=ko.solve_equi(Sigma, m))["s"]
    SiS, Lb = kb.cond_factor(Sigma, s, m=m, ctx=ctx)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    mu = np.zeros(p)
    full = kb.sample(X, mu, SiS, Lb, m=m, Z=None, seed=77, ctx=ctx)
    parts = []
    for r in range(2):
        lo, hi = kb.dist.shard_rows(n, 2, r)
        parts.append(kb.sample(np.asfortranarray(X[lo:hi]), mu, SiS, Lb, m=m, Z=None, seed=77, row_offset=lo, ctx=ctx))
    assert np.array_equal(np.vstack(parts), full)


# ---- group knockoffs ----------------------------------------------------------
def _group_problem(ngroups=10, gsize=5, rho=0.4, gamma=0.2):
    groups = np.repeat(np.arange(1, ngroups + 1), gsize)
    return np.asfortranarray(ko.simulate_block_covariance(groups, rho, gamma)), groups


@pytest.mark.parametrize("method", ["maxent", "mvr", "sdp"])
@pytest.mark.parametrize("pca,ccd", [(1, 1), (1, 0), (10, 5)])
def test_group_solver_matches_oracle_and_reference_properties(kb, ctx, method, pca, ccd):
    """runtests.jl:546-625 configuration (10 groups x 5, rho 0.4, gamma 0.2, m = 5, tol 0.01): parity with the
    oracle on a shared direction set V, plus the reference's own property tests."""
    from oracle import group_oracle as go
    Sigma, groups = _group_problem()
    m = 5
    V = go.get_PCA_vectors(Sigma, list(groups))
    want_S, _, want_obj = go.solve_s_group(Sigma, groups, method, m=m, V=V, tol=0.01, inner_pca_iter=pca,
                                           inner_ccd_iter=ccd)
    keep = Sigma.copy()
    S, gam, obj, info = kb.solve_s_group(Sigma, groups, method, m=m, V=V, tol=0.01, inner_pca_iter=pca,
                                         inner_ccd_iter=ccd, ctx=ctx, return_info=True)
    assert np.array_equal(Sigma, keep)                                  # inputs not mutated (runtests.jl:582-583)
    if method == "sdp" and ccd == 0:
        # PCA-only SDP sweeps end on Brent(abs_tol = 1e-4) of a piecewise-linear objective: its iterate path flips on
        # rounding-level ties, so two correct implementations agree only at the line search's own tolerance.
        assert np.max(np.abs(S - want_S)) < 5e-4 and abs(obj - want_obj) < 1e-3
    else:
        assert np.max(np.abs(S - want_S)) / np.max(np.abs(want_S)) < 1e-8
        assert abs(obj - want_obj) <= 1e-8 * max(1.0, abs(want_obj))
    assert np.linalg.eigvalsh((S + S.T) / 2).min() > -1e-7              # runtests.jl:579-581
    assert np.linalg.eigvalsh((m + 1) / m * Sigma - S).min() > -1e-7
    mask = groups[:, None] == groups[None, :]
    assert np.all(S[~mask] == 0.0)                                      # block diagonal (runtests.jl:584-589)
    assert gam.size == 0


def test_group_solver_default_V_noncontiguous_groups_and_covariance_scale(kb, ctx):
    """groups given out of order (sortperm path, group.jl:370-377, 414-418) on a covariance (cor2cov!, :420);
    the library's own direction set equals the oracle's (per-group eigenvectors)."""
    from oracle import group_oracle as go
    rng = np.random.default_rng(3)
    Sigma, groups = _group_problem(8, 4, 0.5, 0.3)
    p = Sigma.shape[0]
    shuffle = rng.permutation(p)
    d = np.linspace(0.7, 1.9, p)
    Cov = np.asfortranarray((Sigma * np.outer(d, d))[np.ix_(shuffle, shuffle)])
    g2 = groups[shuffle]
    S, _, obj = kb.solve_s_group(Cov, g2, "maxent", m=2, ctx=ctx)
    mask = g2[:, None] == g2[None, :]
    assert np.all(S[~mask] == 0.0)
    assert np.linalg.eigvalsh(1.5 * Cov - S).min() > -1e-7 and np.linalg.eigvalsh(S).min() > -1e-7
    # objective-level agreement with the oracle run on the same problem (direction bases may differ by sign/rotation)
    want_S, _, want_obj = go.solve_s_group(Cov, g2, "maxent", m=2)
    assert abs(obj - want_obj) < 1e-3 * abs(want_obj)
    assert np.max(np.abs(S - want_S)) < 5e-3


def test_group_singletons_fall_back_to_single_solver(kb, ctx):
    Sigma = toeplitz(60, 0.4)
    groups = np.arange(1, 61)
    S, _, obj = kb.solve_s_group(Sigma, groups, "mvr", m=1, ctx=ctx)
    s = kb.solve_s(Sigma, "mvr", m=1, ctx=ctx)
    assert np.array_equal(np.diag(S), s) and obj == 0.0 and np.count_nonzero(S - np.diag(s)) == 0


def test_modelX_group_one_shot(kb, ctx):
    from oracle import group_oracle as go
    rng = np.random.default_rng(4)
    Sigma, groups = _group_problem(6, 5, 0.5, 0.2)
    p, n, m = Sigma.shape[0], 300, 2
    X = np.asfortranarray(rng.standard_normal((n, p)) @ np.linalg.cholesky(Sigma).T)
    Z = np.asfortranarray(rng.standard_normal((n, m * p)))
    V = go.get_PCA_vectors(Sigma, list(groups))
    res = kb.modelX_gaussian_group_knockoffs(X, "maxent", groups, np.zeros(p), Sigma, m=m, V=V, Z=Z, ctx=ctx)
    want_S, _, want_obj = go.solve_s_group(Sigma, groups, "maxent", m=m, V=V)
    want_X = ko.condition(X, np.zeros(p), Sigma, want_S, m=m, Z=Z)
    assert np.max(np.abs(res.S - want_S)) < 1e-8 and abs(res.obj - want_obj) < 1e-8 * abs(want_obj)
    assert np.max(np.abs(res.Xko - want_X)) < 1e-8
    assert res.Xko.shape == (n, m * p)
    with pytest.raises(kb.KnockoffsError):                              # group.jl:119-120
        kb.modelX_gaussian_group_knockoffs(X, "maxent", groups[:-1], np.zeros(p), Sigma, ctx=ctx)


def _objective(Sigma, s, m, method):
    # definitions of group_block_objective (group.jl:24, :27) and utilities.jl:79 on the correlation scale
    kap = (m + 1) / m
    A = kap * Sigma - np.diag(s)
    if method == "sdp_ccd":
        return float(np.sum(s))
    if method == "maxent":
        return float(np.linalg.slogdet(A + 1e-8 * np.eye(len(s)))[1] + m * np.sum(np.log(s + 1e-8)))
    return float(m * m * np.sum(1.0 / s) + np.trace(np.linalg.inv(A)))


@pytest.mark.parametrize("kind,p,method,m", [("ar1", 500, "mvr", 1), ("ar1", 500, "maxent", 1), ("ar1", 300, "sdp_ccd", 1),
                                             ("block", 400, "mvr", 5), ("block", 400, "maxent", 5), ("block", 250, "sdp_ccd", 5)])
@pytest.mark.parametrize("mode", ["block", "stream"])
def test_single_knockoff_objective_values(kb, ctx, kind, p, method, m, mode):
    """north_star: 'Objective values (MVR trace / ME logdet / SDP sum) must agree to 1e-8' -- kb_solve_info.objective
    against the oracle's s pushed through the reference's own objective definitions."""
    Sigma = _sigma(kind, p)
    s0 = ko.solve_equi(Sigma, m)
    want = c_oracle.solve_ccd(Sigma, method, m=m, s_init=s0)
    got, info = kb.solve_s(Sigma, method, m=m, s_init=s0, ctx=ctx, return_info=True, mode=mode)
    obj = _objective(Sigma, want["s"], m, method)
    assert abs(info["objective"] - obj) <= 1e-8 * abs(obj), (info["objective"], obj)
    assert info["flops"] > 0
    _, info0 = kb.solve_s(Sigma, method, m=m, s_init=s0, ctx=ctx, return_info=True, mode=mode, objective=False)
    assert info0["objective"] == 0.0


def test_objective_on_covariance_input_uses_correlation_scale(kb, ctx):
    rng = np.random.default_rng(5)
    p, m = 120, 2
    Sigma = toeplitz(p, 0.4)
    sd = rng.uniform(0.5, 2.0, p)
    Cov = np.asfortranarray(Sigma * np.outer(sd, sd))
    s, info = kb.solve_s(Cov, "maxent", m=m, ctx=ctx, return_info=True)
    obj = _objective(Sigma, s / sd ** 2, m, "maxent")
    assert abs(info["objective"] - obj) <= 1e-8 * abs(obj)


# ---- BASELINE-size cases ---------------------------------------------------------------------------------
def test_config2_maxent_p2000_matches_oracle(kb, ctx):
    """BASELINE configs[1]: AR(1) rho = 0.5, p = 2000, :maxent, m = 1 -- full solve against the C oracle."""
    Sigma = kb.harness.ar1_sigma(2000, 0.5)
    s0 = ko.solve_equi(Sigma, 1)
    want = c_oracle.solve_ccd(Sigma, "maxent", m=1, s_init=s0, lapack_init=True)
    got, info = kb.solve_s(Sigma, "maxent", m=1, s_init=s0, ctx=ctx, return_info=True)
    assert info["sweeps"] == want["sweeps"] and rel(got, want["s"]) < S_RTOL
    # K4 anchor (test/compare_knockpy.ipynb cell 9): interior value of the rho = 0.5 ME solution is p-independent
    assert abs(got[1000] - 0.485332) < 2e-6


def test_config3_full_size_properties(kb, ctx):
    """BASELINE configs[2] at full size (block-equicorrelated, p = 5000, m = 5): too large for the oracle's full
    solve in a test, so size-independent properties: feasibility (PSD of kappa*Sigma - S, 0 <= s <= 1), exchangeability
    of the solution under the block symmetry, first-sweep agreement with the oracle on a bounded number of
    coordinates, and the objective ordering MVR(obj at s_mvr) <= MVR(obj at s_equi)."""
    p, m = 5000, 5
    kap = (m + 1) / m
    Sigma = kb.harness.block_sigma(p, 10, 0.5, 0.25)
    s0 = np.full(p, min(1.0, kap * 0.5))                         # lambda_min of this family is 1 - rho = 0.5
    s, info = kb.solve_s(Sigma, "mvr", m=m, s_init=s0, ctx=ctx, return_info=True)
    assert np.all(s > 0) and np.all(s <= 1)
    assert np.linalg.eigvalsh(kap * Sigma - np.diag(s)).min() > -1e-8
    # all variables are exchangeable up to sweep-order effects: at convergence (tol 1e-6) the entries agree to ~1e-5
    assert s.max() - s.min() < 1e-4
    # bounded oracle sample: the first 300 coordinates of sweep 1 (reference arithmetic) vs the GPU's first sweep
    want = c_oracle.solve_ccd(Sigma, "mvr", m=m, s_init=s0, max_coords=300, lapack_init=True)
    got1 = kb.solve_s(Sigma, "mvr", m=m, s_init=s0, niter=1, mode="stream", ctx=ctx)
    assert rel(got1[:300], want["s"][:300]) < S_RTOL
    # The blocked kernel carries ‖M[:, j]‖² through a 64 x 64 Gram recurrence.  The equi start makes A nearly singular
    # (4500 eigenvalues at λmin = 1e-6, entries of M ~ 1e6, squared column norms ~ 1e12), so in sweep 1 the LAST variable
    # of every 10-block sees its norm fall by ~12 orders of magnitude; ccd_block_seq_kernel detects that cancellation
    # (running error scale of G_jj) and recomputes the norm directly from the panel, so sweep 1 holds 1e-8 as well.
    got1b = kb.solve_s(Sigma, "mvr", m=m, s_init=s0, niter=1, mode="block", ctx=ctx)
    assert rel(got1b[:300], want["s"][:300]) < S_RTOL
    assert rel(got1b, got1) < S_RTOL                       # the whole first sweep, blocked vs rank-1 streaming
    s_stream = kb.solve_s(Sigma, "mvr", m=m, s_init=s0, mode="stream", ctx=ctx)
    assert info["mode"] == "block" and rel(s, s_stream) < 1e-10

    def mvr_obj(sv):
        # m^2 tr(S^-1) + tr((kappa Sigma - S)^-1)   (group.jl:27)
        return m * m * np.sum(1.0 / sv) + np.trace(np.linalg.inv(kap * Sigma - np.diag(sv)))
    assert mvr_obj(s) < mvr_obj(s0 * 0.999)


def test_gram_on_dosage_data(kb, ctx):
    """config-5 style input: 0/1/2 dosages (integers in FP64): the Gram of integers is exact up to the final division."""
    X = kb.harness.dosage_X(6000, 300, 20260005)
    G, mu = kb.gram(X, center=False, denom=1.0, ctx=ctx)
    assert np.array_equal(G, X.T @ X)                          # integer sums < 2^53: bit-exact
    Gc, mu = kb.gram(X, center=True, ctx=ctx)
    Xc = X - X.mean(0)
    assert np.max(np.abs(Gc - Xc.T @ Xc / X.shape[0])) < 1e-12


def test_ledoit_wolf_shrinkage_and_two_argument_entry(kb, ctx):
    """2-argument entry (modelX.jl:39-51): LW-shrunk covariance (row N1) + column means on the device."""
    rng = np.random.default_rng(21)
    n, p = 150, 200                                            # p > n: the case the reference's docs motivate
    Sigma = toeplitz(p, 0.5)
    X = np.asfortranarray(rng.standard_normal((n, p)) @ np.linalg.cholesky(Sigma).T + rng.standard_normal(p))
    want_S, want_mu, want_lam = ko.cov_shrink_lw(X)
    S, mu, lam = kb.cov_shrink_lw(X, ctx=ctx)
    assert abs(lam - want_lam) < 1e-10 and 0 < lam < 1
    assert np.max(np.abs(S - want_S)) < 1e-10 and np.max(np.abs(mu - want_mu)) < 1e-12
    assert np.linalg.eigvalsh(S).min() > 0                      # the point of shrinking (runtests.jl:118-154)
    Z = np.asfortranarray(rng.standard_normal((n, p)))
    res = kb.modelX_gaussian_knockoffs(X, "maxent", Z=Z, ctx=ctx)
    want_X, want_s = ko.modelX_gaussian_knockoffs(X, "maxent", want_mu, want_S, Z=Z)
    assert rel(res.s, want_s) < 1e-7 and np.max(np.abs(res.Xko - want_X)) < 1e-7
    assert np.all(res.s >= 0) and np.linalg.eigvalsh(2 * res.Sigma - np.diag(res.s)).min() > -1e-8


def _blockdiag(mats):
    n = sum(a.shape[0] for a in mats)
    out = np.zeros((n, n), order="F")
    o = 0
    for a in mats:
        out[o:o + a.shape[0], o:o + a.shape[0]] = a
        o += a.shape[0]
    return out


@pytest.mark.parametrize("method", ["maxent", "mvr", "sdp"])
def test_sharded_group_solve_equals_full_solve(kb, method):
    """BASELINE configs[3] logic: block-diagonal Sigma, one LD block per rank, sweeps in lockstep with the reduce hook
    (global equi gamma, global convergence test).  Two ranks are emulated by two host threads with their own contexts
    on this GPU; the union of their S blocks must equal the single-context solve of the whole matrix."""
    import threading
    from oracle import group_oracle as go
    (S1, g1), (S2, g2) = _group_problem(6, 5, 0.4, 0.2), _group_problem(5, 4, 0.6, 0.3)
    full = _blockdiag([S1, S2])
    groups_full = np.concatenate([g1, g2 + g1.max()])
    V1, V2 = go.get_PCA_vectors(S1, list(g1)), go.get_PCA_vectors(S2, list(g2))
    Vfull = _blockdiag_rect(V1, V2)
    m = 3
    # :sdp stops on |Δobj/obj| < tol of a piecewise-linear objective: run it to a tight tolerance so both runs reach the
    # optimum's plateau instead of stopping tol apart
    tol = 1e-6 if method == "sdp" else 1e-3
    want_S, _, want_obj = kb.solve_s_group(full, groups_full, method, m=m, V=Vfull, tol=tol, ctx=kb.Context(0))
    slots = [None, None]
    bar = threading.Barrier(2)
    out = [None, None]

    def worker(rank, Sig, grp, V):
        def reduce(op, vals):
            slots[rank] = (op, vals.copy())
            bar.wait(timeout=60)                     # a rank that dies breaks the barrier instead of hanging the other
            assert slots[0][0] == slots[1][0] and slots[0][1].shape == slots[1][1].shape, "ranks out of lockstep"
            both = np.stack([slots[0][1], slots[1][1]])
            res = both.min(0) if op == 0 else (both.sum(0) if op == 1 else both.max(0))
            bar.wait(timeout=60)
            vals[:] = res
        try:
            out[rank] = kb.solve_s_group_sharded(Sig, grp, method, reduce, m=m, V=V, tol=tol, ctx=kb.Context(0))
        except Exception as e:                       # noqa: BLE001
            out[rank] = e
            bar.abort()

    th = [threading.Thread(target=worker, args=(0, S1, g1, V1), daemon=True),
          threading.Thread(target=worker, args=(1, S2, g2, V2), daemon=True)]
    for t in th:
        t.start()
    for t in th:
        t.join(180)
    for o in out:
        assert o is not None, "rank did not finish"
        if isinstance(o, Exception):
            raise o
    p1 = S1.shape[0]
    # :sdp ends its PCA sweeps on Brent(abs_tol = 1e-4) of a piecewise-linear objective, whose iterate path flips on
    # rounding-level ties (column norms are summed over p vs p_block entries): agreement only at that tolerance
    tol_S, tol_obj = (None, 5e-3) if method == "sdp" else (1e-9, 1e-8 * max(1.0, abs(want_obj)))
    if tol_S is not None:
        assert np.max(np.abs(out[0][0] - want_S[:p1, :p1])) < tol_S
        assert np.max(np.abs(out[1][0] - want_S[p1:, p1:])) < tol_S
    else:
        # the SDP objective sum_g |Sigma_gg - S_gg|_1 / g^2 is flat around its optimum: two iterate paths end on different
        # S of (nearly) the same objective.  What must hold for both: feasibility (runtests.jl:579-589) and the objective.
        S_sh = _blockdiag([out[0][0], out[1][0]])
        kap = (m + 1) / m
        for S_ in (S_sh, want_S):
            # feasible up to the solver's own 1-D step margin eps = 1e-6 (group.jl:1176-1177)
            assert np.linalg.eigvalsh(S_).min() > -1e-6 and np.linalg.eigvalsh(kap * full - S_).min() > -1e-6
    assert abs(out[0][1] + out[1][1] - want_obj) < tol_obj
    assert out[0][2]["inner_sweeps"] == out[1][2]["inner_sweeps"]


def _blockdiag_rect(V1, V2):
    out = np.zeros((V1.shape[0] + V2.shape[0], V1.shape[1] + V2.shape[1]), order="F")
    out[:V1.shape[0], :V1.shape[1]] = V1
    out[V1.shape[0]:, V1.shape[1]:] = V2
    return out


def test_pinned_and_pageable_host_buffers_give_identical_results(kb, ctx):
    """The host-pointer entry points detect PAGEABLE matrices (numpy / Julia arrays) and route them through pinned bounce
    buffers; page-locked buffers are copied directly.  Same bits either way, ragged last chunk and m > 1 included."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(31)
    n, p, m = 9001, 96, 2
    Sigma = toeplitz(p, 0.4)
    X = np.asfortranarray(rng.standard_normal((n, p)) @ np.linalg.cholesky(Sigma).T)
    mu = rng.standard_normal(p) * 0.1
    s = kb.solve_s(Sigma, "maxent", m=m, ctx=ctx)
    Z = np.asfortranarray(rng.standard_normal((n, m * p)))
    ctx.set_sample_chunk_rows(2048)
    try:
        outs = []
        for pinned in (False, True):
            if pinned:
                Xt = torch.empty((p, n), dtype=torch.float64, pin_memory=True); Xt.numpy()[:] = X.T
                Zt = torch.empty((m * p, n), dtype=torch.float64, pin_memory=True); Zt.numpy()[:] = Z.T
                Ot = torch.empty((m * p, n), dtype=torch.float64, pin_memory=True)
                xp, zp, op = Xt.data_ptr(), Zt.data_ptr(), Ot.data_ptr()
            else:
                O = np.empty((n, m * p), order="F")
                xp, zp, op = X.ctypes.data, Z.ctypes.data, O.ctypes.data
            for zz in (zp, None):
                rc = ctx._lib.kb_condition(ctx.h, xp, n, p, mu.ctypes.data, Sigma.ctypes.data, s.ctypes.data, 1, m, zz, 5, 0, op)
                kb.api.check(ctx.h, rc)
                outs.append(Ot.numpy().T.copy() if pinned else O.copy())
        assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[1], outs[3])
        want = ko.condition(X, mu, Sigma, s, m=m, Z=Z)
        assert np.max(np.abs(outs[0] - want)) < XKO_TOL
    finally:
        ctx.set_sample_chunk_rows(0)


def test_tma_staged_sampling_kernel_matches_ldgsts_kernel(kb, ctx):
    """The batched structured sampling launch exists in two forms: LDGSTS / padded rows (default) and TMA
    (cp.async.bulk.tensor + mbarrier ring, 128-byte swizzle, csrc/dgemm_tma.cu; KB_SAMPLE_TMA=1).  Same X-tilde."""
    import os
    rng = np.random.default_rng(41)
    n, p, m = 700, 192, 3
    Sigma = toeplitz(p, 0.5)
    X = np.asfortranarray(rng.standard_normal((n, p)) @ np.linalg.cholesky(Sigma).T)
    Z = np.asfortranarray(rng.standard_normal((n, m * p)))
    mu = rng.standard_normal(p) * 0.1
    s = kb.solve_s(Sigma, "mvr", m=m, ctx=ctx)
    a = kb.condition(X, mu, Sigma, s, m=m, Z=Z, ctx=ctx)
    os.environ["KB_SAMPLE_TMA"] = "1"
    try:
        b = kb.condition(X, mu, Sigma, s, m=m, Z=Z, ctx=ctx)
    finally:
        del os.environ["KB_SAMPLE_TMA"]
    want = ko.condition(X, mu, Sigma, s, m=m, Z=Z)
    assert np.max(np.abs(b - want)) < XKO_TOL and np.max(np.abs(a - want)) < XKO_TOL
    assert np.max(np.abs(a - b)) < 1e-12
