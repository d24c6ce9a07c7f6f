#!/bin/bash
# A/B of dgemm_dmma_kernel build variants (profiles/r01_gemm_variants.md).
#   CPU box:  bash tests/gpu_ab_variants.sh build "16 3 3" "32 2 3" cfg6412     -> build/ab/lib_<variant>.so
#   GPU box:  bash tests/gpu_ab_variants.sh run                                  -> sampling driver + plain shapes per variant
# A variant "BK STAGES CTAS" sets -DKB_GEMM_BK / -DKB_GEMM_STAGES / -DKB_GEMM_MINCTAS; "cfg6412" / "cfg128" set -DKB_GEMM_CFG.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/knockoffs.jl_b200/csrc
mkdir -p $ROOT/build/ab
if [ "$1" = build ]; then
  shift
  make -C $ROOT -s -j8
  for v in "$@"; do
    if [[ "$v" == cfg* ]]; then flags="-DKB_GEMM_CFG=${v#cfg}"; name=$v
    else set -- $v; flags="-DKB_GEMM_BK=$1 -DKB_GEMM_STAGES=$2 -DKB_GEMM_MINCTAS=$3"; name="$1_$2_$3"; fi
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags -c -o /tmp/kb_ab_linalg.o $SRC/linalg.cu
    objs=$(ls $SRC/*.o | grep -v linalg.o)
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build/ab/lib_$name.so $objs /tmp/kb_ab_linalg.o -cudart static
    echo "built build/ab/lib_$name.so ($flags)"
  done
else
  cd $ROOT
  for lib in default build/ab/lib_*.so; do
    if [ "$lib" = default ]; then unset KB_LIB_PATH; else export KB_LIB_PATH=$ROOT/$lib; fi
    echo "== $lib"
    REPS=3 N=16384 python tests/gpu_prof_sampling.py | tail -1
    python tests/gpu_gemm_square.py | head -2
  done
fi
