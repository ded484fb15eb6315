# last validation of the round: GPU suite, RetinaFace stem occupancy A/B (2 / 3 / 4 CTAs per SM), smoke, bench line
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
for occ in 2 3 4 3 2; do
  echo "-- TRB_STEM_OCC=$occ"
  TRB_STEM_OCC=$occ python scripts/profile_ops.py retinaface 2>&1 | grep -E "^==|^ +0 stem" | cut -c1-170
done
TRB_STEM_OCC=4 timeout 300 python -m pytest tests/test_gpu_nets.py -m gpu -q -x -k "retinaface or detection" 2>&1 | grep -E "^E  |passed|failed" | head -4
TRB_STEM_OCC=2 timeout 300 python -m pytest tests/test_gpu_nets.py -m gpu -q -x -k "retinaface or detection" 2>&1 | grep -E "^E  |passed|failed" | head -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -c 12000 gpurun_out/r2j_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['windows_frames_per_s'],'frac',d['roofline']['frac'],d['roofline']['openpose_net_back_to_back'],'launches',d['gpu_launches'],'clocks',d['clocks'])
print(json.dumps(d['per_config']['retinaface_1080p_b32'])[:400])
" 2>&1 | tail -4
