python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-4} --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus ${NG:-4} --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('n',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'h2d',round(d['e2e']['h2d_gbs_per_gpu'],1),'ceiling',round(d['e2e']['h2d_ceiling_gbs_per_gpu'],1),'windows',d['e2e'].get('windows_frames_per_s'),'binding',d['config']['host_binding'])
"
