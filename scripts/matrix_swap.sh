#!/bin/bash
# A/B of conv_tc plan-selection switches inside ONE gpurun call (the pool's boxes differ by ~12 %).
# Usage: bash scripts/matrix_swap.sh VAR "v1 v2 ..." [bench_conv.py filters...]
VAR=${1:-TRB_TC_ISSUERS}; VALS=${2:-"1 0"}; shift 2
for v in $VALS; do echo "== $VAR=$v"; env $VAR=$v python scripts/bench_conv.py "$@"; done
for v in $VALS $VALS; do echo "== nets $VAR=$v"; env $VAR=$v python scripts/profile_ops.py openpose arcface retinaface --brief | grep -E "^==|tcgen05|rror"; done
