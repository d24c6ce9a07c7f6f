"""Tuning aid (not a test): blocked CCD solve time at BASELINE configs[2] for the current KB_CCD_OUTER (inner blocks per
outer block; read once per process).  Usage: KB_CCD_OUTER=k python tests/gpu_outer_sweep.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa: F401,E402

kb = sys.modules["knockoffs_b200"]
HERE = os.path.dirname(os.path.abspath(__file__))
ctx = kb.Context(0)
Sigma = kb.harness.block_sigma(5000, 10, 0.5, 0.25)
out = {"KB_CCD_OUTER": os.environ.get("KB_CCD_OUTER", "4")}
for job in ("c3_sdp_ccd", "c3_mvr", "c3_maxent"):
    z = np.load(os.path.join(HERE, "golden", f"config_{job}.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    best = None
    for _ in range(3):
        t = time.perf_counter()
        s, info = kb.solve_s(Sigma, meta["method"], m=meta["m"], ctx=ctx, return_info=True)
        w = time.perf_counter() - t
        if best is None or w < best[0]:
            best = (w, info)
    err = float(np.max(np.abs(s - z["s"]) / np.abs(z["s"])))
    out[job] = dict(wall_s=best[0], ccd_ms=best[1]["solve_ms"], init_ms=best[1]["init_ms"], sweeps=best[1]["sweeps"],
                    sweep_ms=best[1]["solve_ms"] / best[1]["sweeps"], rel_err=err)
print(json.dumps(out))
