# session Z: read-back epilogue (no TMA store per chunk), merged OpenPose last layers
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
python scripts/bench_patch.py epi 2>&1 | grep -E "dbg=0|dbg=64 "
python scripts/profile_ops.py openpose arcface --brief 2>&1 | grep -E "^==|tcgen05" | cut -c1-150
for r in 1 0; do
TRB_TC_RESIDENT=$r python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench$r.json 2> gpurun_out/r2z_bench.err; python - $r <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r2z_bench{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('RESIDENT',sys.argv[1],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['in_step'])
for k,v in d['per_config'].items(): print(k, json.dumps(v)[:300])
PY
done
