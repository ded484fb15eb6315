b() { echo "-- $*"; env "$@" python scripts/bench_conv.py 'full res' 'arcface 3x3 64' 2>&1 | tail -2; }
b TRB_TC_RESIDENT=1
b TRB_TC_RESIDENT=1 TRB_TC_DEBUG=2
b TRB_TC_RESIDENT=1 TRB_TC_DEBUG=4
b TRB_TC_RESIDENT=1 TRB_TC_DEBUG=8
b TRB_TC_RESIDENT=1 TRB_TC_DEBUG=10
b TRB_TC_RESIDENT=2 TRB_TC_DEBUG=8
