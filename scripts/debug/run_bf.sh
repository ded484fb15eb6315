for v in 1 2; do echo "-- TRB_PATCH=$v"; TRB_PATCH=$v python scripts/profile_ops.py openpose 2>&1 | grep -E "^==|->  57|128-> 256 out  23x 40|->1024" | cut -c1-110 | head -12; done
