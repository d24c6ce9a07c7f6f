"""Diagnostic: group solver, per-group delayed updates (default) vs per-step kernels (KB_GROUP_MODE=step is read once per
process, so this script is run twice) vs the oracle on the PCA-only SDP configuration."""
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
out = {"mode": os.environ.get("KB_GROUP_MODE", "runs")}
for method, pca, ccd in (("sdp", 1, 0), ("sdp", 1, 1), ("sdp", 10, 5), ("maxent", 1, 0), ("mvr", 1, 0)):
    want_S, _, want_obj = go.solve_s_group(Sigma, groups, method, m=m, V=V, tol=0.01, inner_pca_iter=pca, inner_ccd_iter=ccd)
    S, gam, obj, info = kb.solve_s_group(Sigma, groups, method, m=m, V=V, tol=0.01, inner_pca_iter=pca, inner_ccd_iter=ccd,
                                         ctx=ctx, return_info=True)
    rec = go.group_block_objective(Sigma, S, list(groups), m, method)
    rec_w = go.group_block_objective(Sigma, want_S, list(groups), m, method)
    out[f"{method}_{pca}_{ccd}"] = {"maxdiff_S": float(np.max(np.abs(S - want_S))), "obj": obj, "want_obj": want_obj,
                                    "recomputed_obj": rec, "recomputed_want": rec_w, "sweeps": info["inner_sweeps"],
                                    "trace": [round(x, 6) for x in info["trace_obj"][:info["inner_sweeps"]]][:12]}
print(json.dumps(out))
