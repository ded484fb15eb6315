python scripts/debug/c5_timing.py 2>&1 | tail -8
