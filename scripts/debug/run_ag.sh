python scripts/bench_patch.py epi2 2>&1 | tail -16
