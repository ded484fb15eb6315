# final validation of the round on one B200: GPU suite, smoke, the bench line
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -c 12000 gpurun_out/r2h_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['h2d_gbs_per_gpu'],d['e2e']['windows_frames_per_s'],'ceiling',d['e2e']['h2d_ceiling_gbs_per_gpu'],'frac',d['roofline']['frac'],d['roofline']['in_step'],d['roofline']['openpose_net_back_to_back'],'launches',d['gpu_launches'],'cpu',d.get('cpu_baseline',{}).get('value'),'clocks',d['clocks'])
for k,v in d['per_config'].items(): print(k, json.dumps(v)[:420])
" 2>&1 | tail -10
