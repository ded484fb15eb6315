timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -c 9000 gpurun_out/r2f_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['h2d_gbs_per_gpu'],'frac',d['roofline']['frac'],d['roofline']['in_step'],d['roofline']['openpose_net_back_to_back'],'launches',d['gpu_launches'],'cpu',d.get('cpu_baseline',{}).get('value'))
for k,v in d['per_config'].items(): print(k, json.dumps(v)[:420])
" 2>&1 | tail -10
python scripts/profile_ops.py retinaface openpose arcface > gpurun_out/r2f_per_op.txt 2>&1; grep -E "^==|tcgen05" gpurun_out/r2f_per_op.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-per-config > gpurun_out/r2f_ncu_list.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_tc|conv_patch" -c 400 --csv --log-file gpurun_out/r2f_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-per-config > gpurun_out/r2f_ncu_traffic.log 2>&1
ncu --set full --clock-control none -k regex:"conv_patch|conv_tc" -c 53 --csv --page raw --log-file gpurun_out/r2f_ncu_full_convs.csv python scripts/profile_ops.py openpose --brief > gpurun_out/r2f_ncu_full.log 2>&1
