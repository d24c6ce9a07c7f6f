import sys; sys.path.insert(0, ".")
import knockoffs_b200 as kb
S = kb.harness.block_sigma(5000, 10, 0.5, 0.25)
c = kb.Context(0)
kb.solve_s(S, "sdp_ccd", m=5, ctx=c, niter=2)
