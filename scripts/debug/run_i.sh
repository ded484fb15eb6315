timeout 600 python -m pytest tests/test_gpu_conv_patch.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
python scripts/bench_patch.py sk 2>&1 | grep -E "sk=2|trace|launch  [3-6]"
for v in 1; do echo "== TRB_PATCH=$v"; TRB_PATCH=$v python scripts/profile_ops.py openpose --brief 2>&1 | grep -E "^==|conv \*|tcgen05"; done
for v in 1; do echo "== bench TRB_PATCH=$v"; TRB_PATCH=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-per-config 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'people', d['config'].get('people_per_frame'), 'faces', d['config'].get('faces_per_frame'))"; done
