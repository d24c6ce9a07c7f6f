"""B200 profiling driver: the group solver on ONE LD block of BASELINE configs[3] (p = 1250, 250 groups of 5, :mvr, m = 1):
per-group delayed updates = group_panel_kernel -> group_run_seq_kernel -> group_run_apply_kernel per group and sweep.
Prints the solve time and the roofline quantities of group_run_apply_kernel (the kernel that streams D^-1)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa
kb = sys.modules["knockoffs_b200"]
pb = int(os.environ.get("PB", 1250))
g = np.repeat(np.arange(1, pb // 5 + 1), 5)
S = np.where(g[:, None] == g[None, :], 0.75, 0.75 * 0.25).astype(float)
np.fill_diagonal(S, 1.0)
S = np.asfortranarray(S)
ctx = kb.Context(0)
kb.solve_s_group(S[:50, :50], g[:50], "mvr", ctx=ctx)            # warm-up
out = {}
for it in range(int(os.environ.get("REPS", 2))):
    l0 = ctx.launches
    t = time.perf_counter()
    Sg, _, obj, info = kb.solve_s_group(S, g, "mvr", ctx=ctx, return_info=True)
    out = {"p": pb, "groups": int(pb // 5), "wall_s": time.perf_counter() - t, "device_ms": info["solve_ms"], "inner_sweeps": info["inner_sweeps"],
           "steps": info["steps"], "accepted": info["accepted"], "launches": ctx.launches - l0, "obj": obj,
           "touched_bytes": info["touched_bytes"], "us_per_step": 1e3 * info["solve_ms"] / max(1.0, info["steps"]),
           "mode": os.environ.get("KB_GROUP_MODE", "runs (default)")}
print(json.dumps(out))
