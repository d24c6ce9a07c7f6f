"""GPU experiment driver (not a test): final s of the blocked vs the streaming CCD kernel at full size."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200 as kb
ctx = kb.Context(0)
for kind, p, method, m in [("block", 5000, "mvr", 5), ("block", 5000, "maxent", 5), ("ar1", 2000, "mvr", 1), ("block", 5000, "sdp_ccd", 5)]:
    Sigma = kb.harness.block_sigma(p, 10, 0.5, 0.25) if kind == "block" else kb.harness.ar1_sigma(p, 0.5)
    out = {}
    for mode in ("stream", "block"):
        t = time.time()
        s, info = kb.solve_s(Sigma, method, m=m, ctx=ctx, return_info=True, mode=mode)
        out[mode] = (s, info, time.time() - t)
    a, b = out["stream"][0], out["block"][0]
    print(f"{kind} p={p} {method} m={m}: sweeps stream/block {out['stream'][1]['sweeps']}/{out['block'][1]['sweeps']} "
          f"rel|s_block - s_stream| = {np.max(np.abs(a - b) / np.abs(a).clip(1e-300)):.2e}  abs {np.max(np.abs(a-b)):.2e} "
          f"wall {out['stream'][2]:.2f}s / {out['block'][2]:.2f}s  ccd ms {out['stream'][1]['solve_ms']:.0f} / {out['block'][1]['solve_ms']:.0f}", flush=True)
