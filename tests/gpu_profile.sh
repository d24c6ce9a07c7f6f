#!/bin/bash
# ncu evidence for the dominant kernels (B200_PROFILING.md recipe).  Reduced row count (two sampling chunks) so ncu's
# memory save/restore stays small; p = 5000 keeps every kernel at its bench shape.  Outputs -> gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${1:-r01b}
CMD="python bench.py --n 8192 --steps 1 --warmup 1 --no-e2e --no-cpu --no-stream"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "sample/" -k regex:dgemm_dmma_kernel -s 12 -c 1 -o gpurun_out/${TAG}_gemm -f $CMD > gpurun_out/${TAG}_ncu_gemm.log 2>&1
$CMD > gpurun_out/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ccd_block_seq_kernel -s 100 -c 1 -o gpurun_out/${TAG}_seq -f $CMD > gpurun_out/${TAG}_ncu_seq.log 2>&1
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-400
tail -2 gpurun_out/${TAG}_ncu_gemm.log; tail -2 gpurun_out/${TAG}_ncu_seq.log; ls -la gpurun_out | tail -8
