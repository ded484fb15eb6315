# session V: what bounds the 64->64 halo layers (debug modes), new pose_peaks kernel
timeout 600 python -m pytest tests/test_gpu_post.py tests/test_gpu_conv_patch.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py -m gpu -q -x -s 2>&1 | grep -E "^E  |passed|failed|openpose maps|identical|C4 maps" | head -30 | cut -c1-300
b() { echo "-- $*"; env "$@" python scripts/bench_conv.py 'full res' 'arcface 3x3 64' 2>&1 | grep -vE "^$" | head -12; }
b X=0
b TRB_TC_DEBUG=32 
b TRB_TC_DEBUG=1
b TRB_TC_DEBUG=2
b TRB_TC_DEBUG=4
b TRB_TC_DEBUG=8
b TRB_TC_ISSUERS=2 TRB_TC_ISSUERS_HALO=1
b TRB_TC_SUB=1
b TRB_TC_SUB=2
b TRB_TC_HALO=0
b TRB_TC_HALO=2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['in_step'])
for k,v in d['per_config'].items(): print(k, json.dumps(v)[:1500])
PY
