# session Y: warp-local staged epilogue of conv_patch
timeout 900 python -m pytest tests/test_gpu_conv_patch.py tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py tests/test_gpu_capi.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
python scripts/bench_patch.py epi 2>&1 | grep -E "dbg=0|dbg=64 "
python scripts/profile_ops.py openpose arcface 2>&1 | grep -E "^==|tcgen05| [1-9] conv|1[0-9] conv|2[0-9] conv|3[0-9] conv|9[0-9] conv|10[0-9] conv" | cut -c1-150
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['in_step'])
for k,v in d['per_config'].items(): print(k, json.dumps(v)[:400])
PY
