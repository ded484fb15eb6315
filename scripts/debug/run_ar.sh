timeout 1200 python -m pytest tests/test_gpu_conv_patch.py tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
python scripts/profile_ops.py openpose arcface --brief 2>&1 | grep -E "^==|tcgen05" | cut -c1-170
python scripts/bench_patch.py trace 2>&1 | grep -E "launch  [6-8]|trace"
