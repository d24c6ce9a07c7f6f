import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import knockoffs_b200 as kb
ctx = kb.Context(0)
S = kb.harness.block_sigma(5000, 10, 0.5, 0.25)
kb.eigmin(S, ctx=ctx)
l0 = ctx.launches
t = time.time(); lam = kb.eigmin(S, ctx=ctx)
print("eigmin p=5000", lam, "wall %.3f s" % (time.time() - t), "launches", ctx.launches - l0, flush=True)
