# after the shared copy stream: pipeline tests + the bench line
timeout 600 python -m pytest tests -m gpu -q -x -k "pipeline or feeder or streaming or recognition or letterbox" 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
python bench.py --steps 20 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -c 12000 gpurun_out/r2i_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['h2d_gbs_per_gpu'],d['e2e']['windows_frames_per_s'],'ceiling',d['e2e']['h2d_ceiling_gbs_per_gpu'],'frac',d['roofline']['frac'],d['roofline']['in_step'],d['roofline']['openpose_net_back_to_back'],'launches',d['gpu_launches'],'cpu',d.get('cpu_baseline',{}).get('value'),'clocks',d['clocks'])
for k,v in d['per_config'].items(): print(k, json.dumps(v)[:300])
" 2>&1 | tail -10
