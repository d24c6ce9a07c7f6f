#!/bin/bash
# A/B of library build variants (knockoffs.jl_b200/libkb_<name>.so) through bench.py's stages (n = 32768 rows)
for v in base "$@"; do
  if [ $v = base ]; then unset KB_LIB_PATH; else export KB_LIB_PATH=$PWD/knockoffs.jl_b200/libkb_$v.so; fi
  python bench.py --n 32768 --steps 2 --warmup 1 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', {k: round(v,1) for k,v in d['stages_ms'].items()}, 'TF', round(d['roofline']['achieved'],2))"
done
