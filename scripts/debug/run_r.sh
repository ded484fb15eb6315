timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
python scripts/profile_ops.py retinaface arcface --brief 2>&1 | grep -E "^==|25088|tcgen05"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['in_step'])
for k,v in d['per_config'].items(): print(k, json.dumps(v)[:300])
PY
