python bench.py --steps 20 --warmup 3 > gpurun_out/r2bd_bench.json 2> gpurun_out/r2bd_bench.err
tail -c 9000 gpurun_out/r2bd_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['in_step'],d['roofline']['openpose_net_back_to_back'],'launches',d['gpu_launches'],'cpu',d.get('cpu_baseline',{}).get('value'))
for k,v in d['per_config'].items(): print(k, json.dumps(v)[:400])
" 2>&1 | tail -10
python scripts/profile_ops.py retinaface openpose arcface > gpurun_out/r2bd_per_op.txt 2>&1; grep -E "^==|tcgen05" gpurun_out/r2bd_per_op.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2bd_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-per-config > gpurun_out/r2bd_ncu_list.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_tc|conv_patch" -c 400 --csv --log-file gpurun_out/r2bd_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-per-config > gpurun_out/r2bd_ncu_traffic.log 2>&1
