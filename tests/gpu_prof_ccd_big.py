"""Profiling aid (not a test): two :sdp_ccd sweeps at BASELINE configs[2] so that ncu can capture ONE rank-256 pass of the blocked
CCD (the NVTX range "ccd_big" is emitted with KB_NVTX=1; the serial order and the plain stream are needed because ncu cannot
profile kernels on a green-context stream):
    KB_GREEN_SMS=0 KB_CCD_OVERLAP=0 KB_NVTX=1 ncu --set full --clock-control none --nvtx --nvtx-include "ccd_big/" \
        --launch-skip 25 --launch-count 1 -o gpurun_out/r02_ccd_big python tests/gpu_prof_ccd_big.py
With KB_BLOCK_PHASES=1 instead it prints the per-phase split of a sweep."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knockoffs_b200 as kb  # noqa: E402

S = kb.harness.block_sigma(5000, 10, 0.5, 0.25)
c = kb.Context(0)
kb.solve_s(S, "sdp_ccd", m=5, ctx=c, niter=2)
