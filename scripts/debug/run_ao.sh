timeout 900 python -m pytest tests/test_gpu_conv_patch.py tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
python scripts/bench_patch.py epi 2>&1 | grep -E "dbg=0|dbg=64 "
python scripts/profile_ops.py openpose arcface --brief 2>&1 | grep -E "^==|tcgen05" | cut -c1-150
