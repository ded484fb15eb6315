timeout 900 python -m pytest tests/test_gpu_conv_patch.py tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py -m gpu -q -x -s 2>&1 | grep -E "^C[0-9]|passed|failed|^E  |Error" | head -20
python scripts/profile_ops.py arcface --brief 2>&1 | grep -E "^==|28x 28|56x 56|tcgen05"
