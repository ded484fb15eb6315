timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "resize" 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
/usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 5 python -m pytest -m gpu -x -q tests/test_gpu_ops.py -k "resize" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | tail -4
python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from terran_b200.frames import resize_short_side
f = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (32, 1080, 1920, 3), dtype=np.uint8)).cuda()
for side in (416, 184):
    for _ in range(3): resize_short_side(f, side)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): resize_short_side(f, side)
    e1.record(); torch.cuda.synchronize()
    print(side, round(e0.elapsed_time(e1) / 20 * 1e3, 1), 'us')
PY
