"""GPU experiment driver (not a test): group :mvr solve at BASELINE config-4 scale (one LD block and more)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200 as kb
ctx = kb.Context(0)
for p in [int(x) for x in os.environ.get("PS", "250,1250").split(",")]:
    groups = np.repeat(np.arange(1, p // 5 + 1), 5)
    g = groups
    Sigma = np.where(g[:, None] == g[None, :], 0.75, 0.75 * 0.25).astype(float)
    np.fill_diagonal(Sigma, 1.0)
    Sigma = np.asfortranarray(Sigma)
    for method in os.environ.get("METHODS", "mvr,maxent,sdp").split(","):
        t = time.time()
        S, _, obj, info = kb.solve_s_group(Sigma, groups, method, m=1, ctx=ctx, return_info=True)
        wall = time.time() - t
        print(f"p={p} {method}: obj {info['obj_init']:.6g} -> {obj:.6g} outer={info['outer_iters']} sweeps={info['inner_sweeps']} "
              f"steps={info['steps']:.0f} accepted={info['accepted']:.0f} solve_ms={info['solve_ms']:.1f} "
              f"us/step={1e3*info['solve_ms']/max(1,info['steps']):.2f} wall={wall:.2f}s "
              f"eigmin(S)={np.linalg.eigvalsh(S).min():.3g} eigmin(D)={np.linalg.eigvalsh(2*Sigma-S).min():.3g}", flush=True)
