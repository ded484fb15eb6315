timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_baseline_sizes.py tests/test_gpu_capi.py tests/test_gpu_conv_patch.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
for d in 0 1; do echo "== TRB_PT_DUAL=$d"; TRB_PT_DUAL=$d python scripts/profile_ops.py openpose --brief 2>&1 | grep -E "^==|conv \*|tcgen05"; done
TRB_PT_DUAL=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-per-config 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['in_step'])"
