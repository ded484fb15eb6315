for k in 20 100; do python bench.py --steps $k --warmup 3 --no-cpu-baseline --no-per-config 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('steps',d['steps'],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'h2d',round(d['e2e']['h2d_gbs_per_gpu'],1),'ceiling',round(d['e2e']['h2d_ceiling_gbs_per_gpu'],1),'frac',round(d['roofline']['frac'],3))
"; done
