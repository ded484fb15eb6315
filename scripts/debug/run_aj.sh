timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
for r in 1 2; do echo "-- TRB_TC_RESIDENT=$r"; TRB_TC_RESIDENT=$r python scripts/bench_conv.py 'full res' 'arcface 3x3 64' 'retina 3x3' 2>&1 | tail -3; done
for r in 1 2; do echo "-- TRB_TC_RESIDENT=$r"; TRB_TC_RESIDENT=$r python scripts/profile_ops.py retinaface openpose arcface --brief 2>&1 | grep -E "^==" | cut -c1-170; done
