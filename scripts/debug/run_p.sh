timeout 900 python -m pytest tests/test_gpu_conv_patch.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
timeout 300 python scripts/bench_patch.py peek 2>&1
