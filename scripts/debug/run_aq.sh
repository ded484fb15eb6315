timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
for v in 0 1; do echo "-- TRB_TC_LEAN=$v"; TRB_TC_LEAN=$v python scripts/profile_ops.py retinaface openpose arcface --brief 2>&1 | grep -E "^==|tcgen05" | cut -c1-170; done
TRB_TC_LEAN=1 python scripts/bench_conv.py 'full res' 'arcface 3x3 64' 'retina 3x3' 'openpose 1x1 512' 2>&1 | tail -4
