python scripts/debug/e2e_host_time.py 2>&1 | tail -3
