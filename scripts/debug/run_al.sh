python scripts/debug/step_breakdown.py 2>&1 | tail -6
