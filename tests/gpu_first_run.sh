#!/bin/bash
# first GPU shake-out: small cases first, each under its own timeout so a hang cannot eat the box
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40
