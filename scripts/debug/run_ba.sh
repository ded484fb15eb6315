# pool fusion in conv_patch
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py tests/test_gpu_capi.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
for v in 0 1; do echo "-- TRB_POOL_FUSE=$v"; TRB_POOL_FUSE=$v python scripts/profile_ops.py openpose 2>&1 | grep -E "^==|tcgen05|^ +[0-9] |^ 1[01] " | cut -c1-110; done
