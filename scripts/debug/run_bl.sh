for v in 1 2 3 4; do echo "-- TRB_PT_SUB=$v"; TRB_PT_SUB=$v python scripts/profile_ops.py openpose --brief 2>&1 | grep -E "^==" | cut -c1-170; done
for v in 0 2; do echo "-- TRB_PT_SK=$v"; TRB_PT_SK=$v python scripts/profile_ops.py openpose --brief 2>&1 | grep -E "^==" | cut -c1-170; done
