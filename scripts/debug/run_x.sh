# session X: resident-filter halo mode (conv_tc), epilogue debug matrix (conv_patch)
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_conv_patch.py tests/test_gpu_nets.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
for r in 0 1; do echo "-- TRB_TC_RESIDENT=$r"; TRB_TC_RESIDENT=$r python scripts/bench_conv.py 'full res' 'arcface 3x3 64' 'retina 3x3' 2>&1 | tail -3; done
python scripts/bench_patch.py epi 2>&1 | tail -24
