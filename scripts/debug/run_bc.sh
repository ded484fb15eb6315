timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
/usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 5 python -m pytest -m gpu -x -q tests/test_gpu_nets.py -k "openpose_maps or fused_max_pool" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | tail -4
