timeout 900 python -m pytest tests/test_gpu_capi.py tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py tests/test_gpu_conv_patch.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed" | head -20 | cut -c1-300
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['in_step'])
for k,v in d['per_config'].items(): print(k, json.dumps(v)[:900])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2o_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    name=r[4].split('(')[0][:60]; 
    try: t=float(r[-1].replace(',',''))
    except: continue
    agg[name][0]+=1; agg[name][1]+=t
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print(f'{v[1]/1e3:10.1f} us {v[0]:5d}x {100*v[1]/tot:5.1f}%  {k}')
PY
