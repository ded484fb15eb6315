# session AB: clock64 accounting of the patch epilogue; pose overflow test; LRU plan cache
timeout 600 python -m pytest tests/test_gpu_post.py tests/test_gpu_nets.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
python scripts/bench_patch.py epit 2>&1 | tail -12
