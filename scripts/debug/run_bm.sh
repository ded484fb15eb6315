timeout 900 python -m pytest tests/test_gpu_conv_patch.py tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
for v in 2 3 4; do echo "-- TRB_PT_SUB=$v"; TRB_PT_SUB=$v python scripts/profile_ops.py arcface --brief 2>&1 | grep -E "^==" | cut -c1-170; done
python scripts/profile_ops.py openpose 2>&1 | grep -E "^==|tcgen05|^ +[0-9]+ conv" | cut -c1-110 | head -32
