python scripts/debug/h2d_concurrent.py 2>&1 | tail -4
