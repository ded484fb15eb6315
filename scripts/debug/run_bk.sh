timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
/usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 5 python -m pytest -m gpu -x -q tests/test_gpu_conv_patch.py tests/test_gpu_nets.py -k "matches_reference or epilogues or arcface" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | tail -4
