timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_conv_patch.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
for v in 4 6; do echo "-- TRB_TC_NPATCH=$v"; TRB_TC_NPATCH=$v python scripts/bench_conv.py 'full res' 'arcface 3x3 64' 2>&1 | tail -2; done
for v in 1 0; do echo "-- TRB_PT_PA16=$v"; TRB_PT_PA16=$v python scripts/profile_ops.py openpose arcface --brief 2>&1 | grep -E "^==" | cut -c1-170; done
