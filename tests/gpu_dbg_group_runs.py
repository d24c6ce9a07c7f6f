"""Diagnostic: group solver, per-group delayed updates (default) vs per-step kernels (KB_GROUP_MODE=step is read once per
process, so this script is run twice) vs the oracle, sweep by sweep, on the PCA-only SDP configuration."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
from oracle import group_oracle as go, knockoffs_oracle as ko
groups = np.repeat(np.arange(1, 11), 5)
Sigma = np.asfortranarray(ko.simulate_block_covariance(groups, 0.4, 0.2))
m = 5
V = go.get_PCA_vectors(Sigma, list(groups))
ctx = kb.Context(0)
out = {"mode": os.environ.get("KB_GROUP_MODE", "runs"), "graph": os.environ.get("KB_GROUP_GRAPH", "1")}
for method, pca, ccd in (("sdp", 1, 0), ("maxent", 1, 0)):
    rows = []
    for k in (1, 2, 3, 4, 6, 10, 100):
        want_S, _, want_obj = go.solve_s_group(Sigma, groups, method, m=m, V=V, tol=0.01, inner_pca_iter=pca, inner_ccd_iter=ccd,
                                               outer_iter=k)
        S, gam, obj, info = kb.solve_s_group(Sigma, groups, method, m=m, V=V, tol=0.01, inner_pca_iter=pca, inner_ccd_iter=ccd,
                                             outer_iter=k, ctx=ctx, return_info=True)
        rows.append({"outer_iter": k, "sweeps": info["inner_sweeps"], "maxdiff_S": float(np.max(np.abs(S - want_S))),
                     "obj": obj, "want_obj": want_obj, "accepted": info["accepted"], "steps": info["steps"],
                     "recomputed": go.group_block_objective(Sigma, S, list(groups), m, method)})
    out[f"{method}_{pca}_{ccd}"] = rows
print(json.dumps(out))
