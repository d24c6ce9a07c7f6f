"""GPU experiment driver (not a test): potrf / eigmin wall times at p = 5000."""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import knockoffs_b200 as kb
ctx = kb.Context(0)
p = int(os.environ.get("P", "5000"))
S = kb.harness.block_sigma(p, 10, 0.5, 0.25)
for name, A in (("block", S), ("ar1", kb.harness.ar1_sigma(p, 0.5))):
    kb.hook_potrf(A, ctx=ctx)
    l0 = ctx.launches; t = time.time(); L = kb.hook_potrf(A, ctx=ctx); dt = time.time() - t
    print(f"{name}: potrf p={p} wall {dt*1e3:.1f} ms (incl. H2D/D2H of 2x{p*p*8/1e6:.0f} MB) launches {ctx.launches-l0} "
          f"err {np.max(np.abs(L @ L.T - A)):.2e}", flush=True)
    kb.eigmin(A, ctx=ctx)
    l0 = ctx.launches; t = time.time(); lam = kb.eigmin(A, ctx=ctx); dt = time.time() - t
    print(f"{name}: eigmin p={p} {lam!r} wall {dt*1e3:.1f} ms launches {ctx.launches-l0} "
          f"(≈ {(ctx.launches-l0)/ (3*((p+127)//128)+1):.1f} factorisations)", flush=True)
