"""Tuning aid (not a test): parity error and time of the blocked CCD solve at BASELINE configs[2] against the committed
goldens for different rebuild cadences of the resident inverse.  Usage: python tests/gpu_refresh_sweep.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200  # noqa: F401,E402  (loader at the repo root)

kb = sys.modules["knockoffs_b200"]
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ctx = kb.Context(0)
    Sigma = kb.harness.block_sigma(5000, 10, 0.5, 0.25)
    out = {}
    for job in ("c3_sdp_ccd", "c3_mvr", "c3_maxent"):
        z = np.load(os.path.join(HERE, "golden", f"config_{job}.npz"))
        meta = json.loads(bytes(z["meta"]).decode())
        for refresh in (-1, 16, 32, 0):
            kw = dict(refresh_every=refresh) if refresh >= 0 else {}
            if refresh == 0:
                kw["mode"] = "block"
            best = 1e9
            for _ in range(2):
                t = time.perf_counter()
                s, info = kb.solve_s(Sigma, meta["method"], m=meta["m"], ctx=ctx, return_info=True, **kw)
                best = min(best, time.perf_counter() - t)
            err = float(np.max(np.abs(s - z["s"]) / np.abs(z["s"])))
            key = {"mvr": "obj_mvr", "maxent": "obj_maxent", "sdp_ccd": "obj_sdp"}[meta["method"]]
            out[f"{job}/refresh={refresh}"] = dict(rel_err=err, sweeps=info["sweeps"], want_sweeps=meta["sweeps"], wall_s=best,
                                                   ccd_ms=info["solve_ms"], init_ms=info["init_ms"], mode=info["mode"],
                                                   obj_rel=abs(info["objective"] - meta[key]) / abs(meta[key]))
            print(job, refresh, out[f"{job}/refresh={refresh}"], flush=True)
    json.dump(out, open("gpurun_out/r02_refresh_sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()
