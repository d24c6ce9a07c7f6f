#!/bin/bash
# ncu evidence for the two dominant kernels (B200_PROFILING.md recipe).  Reduced row count (one sampling
# chunk) so ncu's memory save/restore stays small; p = 5000 keeps both kernels at their bench shapes.
set -x
mkdir -p gpurun_out
CMD="python bench.py --n 8192 --steps 1 --warmup 0 --no-e2e --no-cpu"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ccd_stream_kernel -s 1 -c 1 -o gpurun_out/prof_ccd -f $CMD > gpurun_out/ncu_ccd.log 2>&1
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "sample/" -k regex:dgemm_dmma_kernel -s 1 -c 1 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/prof_plain.log | cut -c1-600
tail -3 gpurun_out/ncu_ccd.log; tail -3 gpurun_out/ncu_gemm.log; ls -la gpurun_out
