b() { echo "-- $*"; env "$@" python scripts/profile_ops.py openpose --brief 2>&1 | grep -E "^==" | cut -c60-170; }
b X=1
b TRB_TC_PDL_LATE=0
b TRB_PT_STAGES=4
b TRB_PT_STAGES=3
b TRB_PT_STAGES=2
