timeout 900 python -m pytest tests/test_gpu_conv_patch.py tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py tests/test_gpu_capi.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
for v in 0 1; do echo "-- TRB_PT_STACK=$v"; TRB_PT_STACK=$v python scripts/profile_ops.py openpose arcface --brief 2>&1 | grep -E "^==|tcgen05" | cut -c1-170; done
TRB_PT_STACK=1 python scripts/profile_ops.py arcface 2>&1 | grep -E " (38|39|40|41|99|100|101|102) conv" | cut -c1-120
