timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -c 9000 gpurun_out/r2g_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['h2d_gbs_per_gpu'],'frac',d['roofline']['frac'],d['roofline']['in_step'],d['roofline']['openpose_net_back_to_back'],'launches',d['gpu_launches'],'cpu',d.get('cpu_baseline',{}).get('value'))
for k,v in d['per_config'].items(): print(k, json.dumps(v)[:420])
" 2>&1 | tail -10
python scripts/profile_ops.py retinaface openpose arcface > gpurun_out/r2g_per_op.txt 2>&1; grep -E "^==|tcgen05" gpurun_out/r2g_per_op.txt
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
