for v in 1 2; do echo "-- TRB_PATCH=$v"; TRB_PATCH=$v python scripts/profile_ops.py openpose arcface 2>&1 | grep -E "^==|64->  64|64-> 128|tcgen05" | cut -c1-120 | head -24; done
