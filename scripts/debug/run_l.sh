timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-300
python scripts/profile_ops.py openpose arcface --brief 2>&1 | grep -E "^==|  64->  64|64-> 128|tcgen05"
python bench.py --steps 20 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -c 3000 gpurun_out/r2n_bench.json
