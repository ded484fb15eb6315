# 1 GPU: the e2e window repeated, feeder on a private stream (old) against the shared copy stream
cat /proc/loadavg
MODES=private,spin,private,spin STEPS=20 WINDOWS=6 python scripts/debug/e2e_ranks.py 2>&1 | grep -E "^==|rank 0|max over|slowest|first result"
