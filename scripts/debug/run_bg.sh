timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv" 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
for v in 16 0; do echo "-- TRB_TC_PW=$v"; TRB_TC_PW=$v python scripts/bench_conv.py 'full res' 'arcface 3x3 64' 'retina 3x3' 2>&1 | tail -3; done
