"""B200: timings of the SURVEY 8(f) rows at representative sizes (N2 approx windows, N3 representative groups,
N4 fused fit_marginal), each beside the CPU oracle on a bounded sample.  Prints one JSON object.
    python tests/gpu_widen_bench.py > gpurun_out/widen_bench.json"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
from oracle import approx_oracle as ao, rep_oracle as ro, stats_oracle as so, knockoffs_oracle as ko  # CPU baseline legs only

ctx = kb.Context(0)
out = {}
rng = np.random.default_rng(20260020)


def timed(f, reps=1):
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter(); r = f(); best = min(best, time.perf_counter() - t)
    return best, r


# ---- N4: fit_marginal fused (host X in, X̃ never leaves the device) at BASELINE configs[2] -----------------
p, n, m = 5000, int(os.environ.get("N4_ROWS", 100000)), 5
Sigma = kb.harness.block_sigma(p, 10, 0.5, 0.25)
L = np.linalg.cholesky(Sigma)
X = np.empty((n, p), order="F")
for r0 in range(0, n, 20000):
    X[r0:r0 + 20000] = rng.standard_normal((min(20000, n - r0), p)) @ L.T
beta = np.zeros(p); idx = rng.choice(p, 50, replace=False); beta[idx] = rng.uniform(0.05, 0.1, 50)
y = X @ beta + rng.standard_normal(n)
mu = np.zeros(p)
kb.fit_marginal(y, X[:4096], mu, Sigma, method="mvr", m=m, keep_knockoffs=False, ctx=ctx) if False else None
t_warm, _ = timed(lambda: kb.fit_marginal(y, X, mu, Sigma, method="mvr", m=m, keep_knockoffs=False, seed=1, ctx=ctx))
t_fused, r = timed(lambda: kb.fit_marginal(y, X, mu, Sigma, method="mvr", m=m, keep_knockoffs=False, seed=1, ctx=ctx))
sel = set(r.selected[2] - 1)
rows_cpu = 2000
Xs = X[:rows_cpu]; Xk = rng.standard_normal((rows_cpu, m * p))
t_cpu, _ = timed(lambda: so.marginal_T(y[:rows_cpu], Xs, Xk, m))
out["N4_fit_marginal_fused"] = {"n": n, "p": p, "m": m, "seconds": t_fused, "rows_per_s": n / t_fused,
                                "h2d_bytes": int((n * p + p * p + n + p) * 8), "d2h_bytes": int(((m + 3) * p + 5 * p) * 8),
                                "selected_at_q0.1": len(sel), "true_positives": len(sel & set(idx)),
                                "cpu_oracle_scores_seconds_extrapolated": t_cpu * n / rows_cpu,
                                "cpu_sample": f"numpy z-score + X'y on {rows_cpu} of {n} rows (x n/rows), all BLAS threads"}
# the same call with X in PAGE-LOCKED host memory (what bench.py's e2e uses); the default above is pageable numpy memory
import torch
Xp_t = torch.empty((p, n), dtype=torch.float64, pin_memory=True)
Xp = Xp_t.numpy().T                                   # (n, p) Fortran-ordered view of the pinned buffer
Xp[:] = X
timed(lambda: kb.fit_marginal(y, Xp, mu, Sigma, method="mvr", m=m, keep_knockoffs=False, seed=1, ctx=ctx))
t_pin, r2 = timed(lambda: kb.fit_marginal(y, Xp, mu, Sigma, method="mvr", m=m, keep_knockoffs=False, seed=1, ctx=ctx))
out["N4_fit_marginal_fused"].update({"host_memory": "pageable (numpy)", "seconds_pinned_X": t_pin, "rows_per_s_pinned_X": n / t_pin,
                                      "same_selection_pinned": bool(np.array_equal(r2.selected[2], r.selected[2]))})
del X, y, Xp, Xp_t

# ---- N2: approx_modelX_gaussian_knockoffs, p = 5000 in windows of 500 --------------------------------------
n2, p2, m2 = 20000, 5000, 1
S2 = kb.harness.ar1_sigma(p2, 0.5)
X2 = np.asfortranarray(rng.standard_normal((n2, p2)) @ np.linalg.cholesky(S2).T)
timed(lambda: kb.approx_modelX_gaussian_knockoffs(X2[:, :1000], "maxent", windowsize=500, m=m2, seed=2, dense_sigma=False, ctx=ctx))
t_gpu, a = timed(lambda: kb.approx_modelX_gaussian_knockoffs(X2, "maxent", windowsize=500, m=m2, seed=2, dense_sigma=False, ctx=ctx))
t_cpu, _ = timed(lambda: ao.approx_modelX_gaussian_knockoffs(X2[:2000, :500], "maxent", [(0, 500)], m=m2,
                                                             Z=np.zeros((2000, 500))))
out["N2_approx_windows"] = {"n": n2, "p": p2, "windowsize": 500, "windows": 10, "method": "maxent", "m": m2, "seconds": t_gpu,
                            "rows_per_s": n2 / t_gpu, "gamma": a.gamma, "sweeps_max": a.info["sweeps"],
                            "cpu_oracle_one_window_2000_rows_seconds": t_cpu,
                            "cpu_sample": "oracle on 1 of 10 windows and 2000 of 20000 rows"}
del X2

# ---- N3: representative group knockoffs, p = 2000 in 400 groups of 5 -----------------------------------------
p3, n3 = 2000, 5000
groups = np.repeat(np.arange(1, p3 // 5 + 1), 5)
S3 = np.asfortranarray(ko.simulate_block_covariance(groups, 0.75, 0.25))
kb.choose_group_reps(S3, groups, threshold=0.5, ctx=ctx)          # warm-up: workspace + module load
t_reps, reps = timed(lambda: kb.choose_group_reps(S3, groups, threshold=0.5, ctx=ctx))
X3 = np.asfortranarray(rng.standard_normal((n3, p3)) @ np.linalg.cholesky(S3).T)
t_all, rr = timed(lambda: kb.modelX_gaussian_rep_group_knockoffs(X3, "maxent", groups, np.zeros(p3), S3, m=1, seed=3,
                                                                 group_reps=reps, ctx=ctx))
ps = 200
t_cpu, reps_cpu = timed(lambda: ro.choose_group_reps(S3[:ps, :ps], groups[:ps], threshold=0.5))
out["N3_rep_group"] = {"p": p3, "n": n3, "groups": int(p3 // 5), "representatives": int(len(reps)),
                       "choose_group_reps_seconds": t_reps, "knockoffs_seconds_given_reps": t_all, "obj": rr.obj,
                       "cpu_oracle_choose_group_reps_p200_seconds": t_cpu,
                       "cpu_sample": f"literal oracle on the leading {ps} x {ps} block (its cost grows like p^3 per group step)"}
print(json.dumps(out))
