"""Tiny end-to-end pass over every entry point added in r01c-r01e, meant to be run under
`compute-sanitizer --tool memcheck` (one tool per gpurun call, see B200_PROFILING.md)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
ctx = kb.Context(0)
rng = np.random.default_rng(0)
p, n, m = 70, 333, 3
S = kb.harness.ar1_sigma(p, 0.5)
X = kb.harness.gaussian_X(n, S, 1)
y = X[:, 0] - X[:, 5] + rng.standard_normal(n)
ctx.set_sample_chunk_rows(128)
r = kb.fit_marginal(y, X, np.zeros(p), S, method="mvr", m=m, seed=1, ctx=ctx)
r1 = kb.fit_marginal(y, X, np.zeros(p), S, method="maxent", m=1, seed=1, keep_knockoffs=False, ctx=ctx)
k = kb.modelX_gaussian_knockoffs(X, "sdp_ccd", np.zeros(p), S, m=2, seed=3, ctx=ctx)
SiS, Lb = kb.cond_factor(S, 0.5 * k.s, m=4, ctx=ctx)
Xk = kb.sample(X, np.zeros(p), SiS, Lb, m=4, seed=5, ctx=ctx)
Lb2 = Lb.copy(); Lb2[3, 1, 1] += 0.5                    # breaks the L + upper structure -> general form
Xk2 = kb.sample(X, np.zeros(p), SiS, Lb2, m=4, seed=5, ctx=ctx)
ctx.set_sample_chunk_rows(0)
a = kb.approx_modelX_gaussian_knockoffs(X, "maxent", [(1, 25), (26, 26), (27, 70)], m=2, seed=2, ctx=ctx)
g = np.repeat(np.arange(1, p // 5 + 1), 5)
reps = kb.choose_group_reps(S, g, threshold=0.6, ctx=ctx)
rr = kb.modelX_gaussian_rep_group_knockoffs(X, "maxent", g, np.zeros(p), S, m=2, seed=4, ctx=ctx)
N = kb.cond_indep_corr(S, g, reps, ctx=ctx)
gk = kb.modelX_gaussian_group_knockoffs(X, "mvr", g, np.zeros(p), S, m=2, seed=6, ctx=ctx)
f = kb.fit_marginal(y, gk, ctx=ctx)
print("ok", r.W.shape, r1.W.shape, Xk.shape, float(np.abs(Xk - Xk2).max()) > 0, a.gamma, len(reps), rr.obj, N.shape, f.W.shape)
