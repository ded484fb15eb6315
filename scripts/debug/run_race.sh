for v in 1 0; do
echo "=== TRB_TC_RESIDENT=$v TRB_TC_LEAN=$v"
TRB_TC_RESIDENT=$v TRB_TC_LEAN=$v timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30 python -m pytest -m gpu -x -q tests/test_gpu_ops.py -k "conv_matches_reference or conv_stream_k" 2>&1 | grep -E "Race reported|Hazard|Error:|Warning:|=========     at |RACECHECK SUMMARY|passed|failed" | sed 's/^=========//' | sort | uniq -c | sort -rn | head -30
done
