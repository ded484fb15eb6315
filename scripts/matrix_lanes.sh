for cfg in "TRB_LANES=1 TRB_TC_PDL_LATE=0" "TRB_LANES=1 TRB_TC_PDL_LATE=1" "TRB_LANES=0 TRB_TC_PDL_LATE=1" "TRB_LANES=1 TRB_TC_PDL=0" "TRB_LANES=0 TRB_TC_PDL_LATE=0"; do
  echo "== $cfg"
  env $cfg python bench.py --no-cpu-baseline --no-per-config 2>&1 | grep '"metric"' | python -c "import json,sys; d=json.loads(sys.stdin.read().strip()); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e', round(d['e2e']['value'],1))"
done
