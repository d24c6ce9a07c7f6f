"""Generates the full-size BASELINE.json config fixtures under tests/golden/ with the C oracle
(oracle/ccd_oracle.c -- the line-by-line restatement of utilities.jl:124-343, 516-625, pinned to K1-K4).

    python tests/golden/make_config_goldens.py [c2_maxent] [c3_mvr] [c3_sdp_ccd] [c5_maxent] ...

Each job is one single-threaded CPU solve (the reference's sweep is sequential); wall time in this 8-core container:
c2_maxent ~20 s, c3_mvr ~6 min, c3_maxent ~8 min, c5_maxent ~10 min, c3_sdp_ccd (59 sweeps) ~70 min.
Fixtures hold only the p-vectors (s, s_init) + sweep count + per-sweep trace + the generator parameters, ~40-80 KB each.
TEST INFRASTRUCTURE: nothing here is imported by the product.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import knockoffs_b200  # noqa: E402,F401  (loader; only harness generators are used, no GPU call)
from oracle import c_oracle, knockoffs_oracle as ko  # noqa: E402

harness = sys.modules["knockoffs_b200"].harness

C5_GOLDEN_ROWS = 20000      # rows of the dosage matrix behind the c5 Sigma-hat fixture (the Gram is exact integer arithmetic)
C5_SEED = 20260005


def c5_sigma_hat(n=C5_GOLDEN_ROWS, p=5000):
    """Sigma-hat = Xc'Xc/n of the dosage generator (modelX.jl:39-51 with the plain covariance), then cov2cor is left to solve_s."""
    X = harness.dosage_X(n, p, C5_SEED)
    mu = X.mean(axis=0)
    Xc = X - mu[None, :]
    S = (Xc.T @ Xc) / n
    return np.asfortranarray(S), mu


def job_sigma(name):
    if name.startswith("c2"):
        return harness.ar1_sigma(2000, 0.5), 1
    if name.startswith("c3"):
        return harness.block_sigma(5000, 10, 0.5, 0.25), 5
    if name.startswith("c5"):
        return c5_sigma_hat()[0], 1
    raise SystemExit(f"unknown job {name}")


def run(name):
    Sigma, m = job_sigma(name)
    method = name.split("_", 1)[1]
    p = Sigma.shape[0]
    # solve_s (utilities.jl:27-51): cov -> cor, solve, s .*= sigma^2
    sd = np.sqrt(np.diag(Sigma))
    cor = Sigma if np.allclose(sd, 1.0) else np.asfortranarray(Sigma / np.outer(sd, sd))
    s_init = None if method == "sdp_ccd" else ko.solve_equi(cor, m)
    t = time.time()
    r = c_oracle.solve_ccd(cor, method, m=m, s_init=s_init)
    wall = time.time() - t
    assert r["rc"] == 0, r["rc"]
    s = r["s"] * sd ** 2
    meta = dict(job=name, p=p, m=m, method=method, sweeps=int(r["sweeps"]), trace=[float(x) for x in r["trace"]],
                wall_s=wall, oracle="oracle/ccd_oracle.c via c_oracle.solve_ccd, defaults niter=100 tol=1e-6 lambda_min=1e-6",
                generator={"c2": "harness.ar1_sigma(2000, 0.5)", "c3": "harness.block_sigma(5000, 10, 0.5, 0.25)",
                           "c5": f"Gram/n of harness.dosage_X({C5_GOLDEN_ROWS}, 5000, {C5_SEED}) centred"}[name[:2]])
    obj = objectives(cor, r["s"], m)
    meta.update(obj)
    np.savez_compressed(os.path.join(HERE, f"config_{name}.npz"), s=s, s_cor=r["s"],
                        s_init=np.zeros(0) if s_init is None else s_init, sd=sd,
                        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    print(name, "sweeps", r["sweeps"], "wall %.1f s" % wall, "s range", s.min(), s.max(), obj, flush=True)


def objectives(cor, s, m):
    """North-star objective values of the single-knockoff solvers, definitions of group.jl:24,27 and utilities.jl:79
    (SURVEY section 8 text under table a): MVR m^2 tr(S^-1) + tr((kappa Sigma - S)^-1); ME logdet(kappa Sigma - S + 1e-8 I)
    + m logdet(S + 1e-8 I); SDP sum(s).  Evaluated on the correlation scale the solver works in."""
    import scipy.linalg as sla
    kap = (m + 1) / m
    A = kap * cor - np.diag(s)
    out = {"obj_sdp": float(np.sum(s))}
    try:
        c = sla.cholesky(A + 1e-8 * np.eye(len(s)), lower=True, check_finite=False)
        out["obj_maxent"] = float(2 * np.sum(np.log(np.diag(c))) + m * np.sum(np.log(s + 1e-8)))
    except Exception:
        out["obj_maxent"] = None
    try:
        c = sla.cholesky(A, lower=True, check_finite=False)
        ci = sla.solve_triangular(c, np.eye(len(s)), lower=True, check_finite=False)
        out["obj_mvr"] = float(m * m * np.sum(1.0 / s) + np.sum(ci * ci))
    except Exception:
        out["obj_mvr"] = None
    return out


if __name__ == "__main__":
    for job in (sys.argv[1:] or ["c2_maxent", "c3_mvr", "c3_maxent", "c5_maxent", "c3_sdp_ccd"]):
        run(job)
