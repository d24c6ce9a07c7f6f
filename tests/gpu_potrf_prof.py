import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import knockoffs_b200 as kb
ctx = kb.Context(0)
rng = np.random.default_rng(0)
B = rng.standard_normal((640, 700)); A = B @ B.T / 640 + np.eye(640)
L = kb.hook_potrf(A, ctx=ctx)
print(np.max(np.abs(L - np.linalg.cholesky(A))))
