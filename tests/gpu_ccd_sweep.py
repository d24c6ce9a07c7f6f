"""GPU experiment driver (not a test): times the CCD kernel at p=5000 under env-selected variants."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200 as kb
p = int(os.environ.get("P", "5000")); m = int(os.environ.get("M", "5"))
Sigma = kb.harness.block_sigma(p, 10, 0.5, 0.25)
s0 = np.full(p, min(1.0, (m + 1) / m * 0.5))
ctx = kb.Context(0)
for rep in range(2):
    t = time.time()
    s, info = kb.solve_s(Sigma, os.environ.get("METHOD", "mvr"), m=m, s_init=s0, ctx=ctx, return_info=True,
                         refresh_every=int(os.environ.get("REFRESH", "-1")), mode=os.environ.get("MODE", "auto"), niter=int(os.environ.get("NITER", "100")))
    wall = time.time() - t
    per = info["solve_ms"] / max(1, info["sweeps"])
    print(f"mode={info['mode']} shape={os.environ.get('KB_CCD_SHAPE','2x256')} l2={os.environ.get('KB_CCD_L2_PERSIST','0')} p={p} sweeps={info['sweeps']} "
          f"ccd_ms/sweep={per:.2f} GB/s={info['ref_bytes']/max(1,info['sweeps'])/per/1e6:.0f} init_ms={info['init_ms']:.1f} wall={wall:.2f}s "
          f"s[:3]={s[:3]}", flush=True)
