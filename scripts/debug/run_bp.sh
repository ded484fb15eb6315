timeout 900 python -m pytest tests/test_gpu_conv_patch.py tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py tests/test_gpu_capi.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
for v in 0 1; do echo "-- TRB_PT_WIDE=$v"; TRB_PT_WIDE=$v python scripts/profile_ops.py openpose arcface --brief 2>&1 | grep -E "^==" | cut -c1-170; done
TRB_PT_WIDE=1 python scripts/profile_ops.py openpose 2>&1 | grep -E "^ +(3|4|6|19|26) conv" | cut -c1-110
