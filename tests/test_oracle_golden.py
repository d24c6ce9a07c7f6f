"""Pins the CPU oracle (oracle/) against the reference's own published known-answer vectors
(tests/golden/reference_known_answers.json).  CPU only."""
import numpy as np
import pytest

from conftest import toeplitz
from oracle import c_oracle, knockoffs_oracle as ko


def _check(s, g):
    tol = g.get("tol", 0.6 * 10.0 ** (-g["digits"]) if g["digits"] < 16 else 1e-13)
    h, t = np.array(g["head"]), np.array(g["tail"])
    assert np.max(np.abs(s[:len(h)] - h)) <= tol
    assert np.max(np.abs(s[len(s) - len(t):] - t)) <= tol


@pytest.mark.parametrize("key,rho", [("K1_mvr_ar1_rho0.4_p1000_m1", 0.4), ("K2_maxent_ar1_rho0.4_p1000_m1", 0.4),
                                     ("K3_maxent_ar1_rho0.4_p1000_m5", 0.4), ("K4_sdp_ccd_ar1_rho0.5_p300_m1", 0.5),
                                     ("K4_mvr_ar1_rho0.5_p300_m1", 0.5), ("K4_maxent_ar1_rho0.5_p300_m1", 0.5)])
def test_c_oracle_matches_reference_docs(golden, key, rho):
    g = golden[key]
    Sigma = toeplitz(g["p"], rho)
    s0 = ko.solve_equi(Sigma, g["m"])
    r = c_oracle.solve_ccd(Sigma, g["method"], m=g["m"], s_init=s0)
    assert r["rc"] == 0
    _check(r["s"], g)


def test_equi_matches_reference_docs(golden):
    g = golden["K2_equi_ar1_rho0.4_p1000_m1"]
    s = ko.solve_equi(toeplitz(1000, 0.4), 1)
    assert abs(s[0] - g["head"][0]) < 1e-6 and np.all(s == s[0])


@pytest.mark.parametrize("method", ["mvr", "maxent", "sdp_ccd"])
def test_numpy_oracle_equals_c_oracle(method):
    """The two independent restatements (numpy/LAPACK solves vs plain C loops) agree to rounding."""
    Sigma = toeplitz(120, 0.5)
    s0 = ko.solve_equi(Sigma, 1)
    r = c_oracle.solve_ccd(Sigma, method, m=1, s_init=s0)
    if method == "mvr":
        s, sw = ko.solve_MVR(Sigma, s_init=s0)
    elif method == "maxent":
        s, sw = ko.solve_max_entropy(Sigma, s_init=s0)
    else:
        s, sw = ko.solve_sdp_ccd(Sigma)
    assert sw == r["sweeps"]
    assert np.max(np.abs(s - r["s"])) < 1e-12


def test_sdp_ccd_sweep_count_and_kwargs():
    """utilities.jl:339-340: lambda 0.5 * 0.8^k < 1e-6 first at k = 59; runtests.jl:349-365 kwargs agree to 0.05."""
    Sigma = toeplitz(100, 0.4)
    a = c_oracle.solve_ccd(Sigma, "sdp_ccd", sdp_lambda=0.7, sdp_mu=0.7)
    b = c_oracle.solve_ccd(Sigma, "sdp_ccd", sdp_lambda=0.9, sdp_mu=0.9)
    assert c_oracle.solve_ccd(Sigma, "sdp_ccd")["sweeps"] == 59
    assert np.max(np.abs(a["s"] - b["s"])) < 0.05


@pytest.mark.parametrize("method,m", [("mvr", 3), ("maxent", 5), ("sdp_ccd", 5)])
def test_oracle_properties_multiple_knockoffs(method, m):
    """runtests.jl:773-808: 0 <= s <= 1 and (m+1)/m Sigma - diag(s) is PSD."""
    Sigma = toeplitz(150, 0.4)
    r = c_oracle.solve_ccd(Sigma, method, m=m, s_init=ko.solve_equi(Sigma, m))
    s = r["s"]
    assert r["rc"] == 0 and np.all(s >= 0) and np.all(s <= 1 + 1e-12)
    assert np.linalg.eigvalsh((m + 1) / m * Sigma - np.diag(s)).min() > -1e-8


def test_rank1_updates_match_refactorisation():
    """lowrank{up,down}date_turbo! (utilities.jl:516-625) equal chol(A +/- vv') -- the check the
    reference's TODO list (runtests.jl:810-818) asks for."""
    rng = np.random.default_rng(0)
    p = 60
    B = rng.standard_normal((p, p))
    A = B @ B.T + p * np.eye(p)
    for j, d in [(0, 0.7), (17, 3.0), (p - 1, 0.1)]:
        U = np.ascontiguousarray(np.linalg.cholesky(A).T)
        v = np.zeros(p); v[j] = np.sqrt(d)
        ko.lowrankupdate_turbo(U, v.copy())
        A2 = A.copy(); A2[j, j] += d
        assert np.max(np.abs(U - np.linalg.cholesky(A2).T)) < 1e-12
        ko.lowrankdowndate_turbo(U, v.copy())
        assert np.max(np.abs(U - np.linalg.cholesky(A).T)) < 1e-12


def test_downdate_raises_posdef():
    U = np.eye(3)
    v = np.array([0.0, 1.5, 0.0])
    with pytest.raises(np.linalg.LinAlgError):
        ko.lowrankdowndate_turbo(U, v)


def test_condition_block_structure_equals_dense_factor():
    """the pm x pm factor of modelX.jl:116-120 has L_kk / M_k block structure."""
    rng = np.random.default_rng(1)
    p, m = 30, 4
    Sigma = toeplitz(p, 0.5)
    s = c_oracle.solve_ccd(Sigma, "maxent", m=m, s_init=ko.solve_equi(Sigma, m))["s"]
    SiS, L = ko.cond_factor(Sigma, s, m)
    for k in range(m):
        for i in range(k + 1, m):
            assert np.max(np.abs(L[i * p:(i + 1) * p, k * p:(k + 1) * p] - L[(k + 1) * p:(k + 2) * p, k * p:(k + 1) * p])) < 1e-13
    X = rng.standard_normal((20, p)); Z = rng.standard_normal((20, m * p))
    Xko = ko.condition(X, np.zeros(p), Sigma, s, m=m, Z=Z)
    assert Xko.shape == (20, m * p)


def test_golden_fixture_matches_its_sources():
    """tests/golden/make_golden.py re-reads the cited lines of the reference checkout (build container only; the GPU box has
    no /root/reference) and checks every number of the fixture against them."""
    import os
    import subprocess
    import sys
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py")
    r = subprocess.run([sys.executable, script, ref], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
