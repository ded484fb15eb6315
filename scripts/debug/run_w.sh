# session W: staging-ring epilogue of conv_patch, new pose kernels
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
python scripts/profile_ops.py openpose arcface 2>&1 | grep -E "^==|conv|tcgen05" | cut -c1-150
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['in_step'])
for k,v in d['per_config'].items(): print(k, json.dumps(v)[:1500])
PY
