python scripts/debug/feeder_timeline.py 2>&1 | tail -34
