for v in 0 1; do echo "-- TRB_FPN_MMA=$v"; TRB_FPN_MMA=$v python scripts/profile_ops.py retinaface 2>&1 | grep -E "^==|^ 1[4-8] " | cut -c1-150; done
TRB_FPN_MMA=1 timeout 600 python -m pytest tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py -m gpu -q -x -k "retina or detection or c2" 2>&1 | grep -E "^E  |passed|failed" | head -5
