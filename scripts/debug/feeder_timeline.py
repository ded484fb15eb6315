"""Copy-stream timeline of the e2e pipeline: per-batch H2D duration and the idle gap before it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('TERRAN_HOME', os.path.join(ROOT, '.pytest_cache', 'terran_home'))
os.makedirs(os.environ['TERRAN_HOME'], exist_ok=True)
import numpy as np, torch
import bench
from terran_b200 import pipeline as pl
from terran_b200.face.detection import Detection
from terran_b200.face.detection.retinaface import RetinaFace
from terran_b200.pose import Estimation
from terran_b200.pose.openpose import OpenPose
dev = torch.device('cuda')
sd_det, sd_pose = bench.bench_weights()
det = Detection(device=dev, lazy=True); det.model = RetinaFace(device=dev, state_dict=sd_det)
est = Estimation(device=dev, lazy=True); est.model = OpenPose(device=dev, state_dict=sd_pose)
host = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (32, 1080, 1920, 3), dtype=np.uint8)).pin_memory()
pipe = pl.PerceptionPipeline(det, est, device=dev)
for _ in pipe.run(pl.FrameFeeder((host for _ in range(3)), device=dev)): pass
events = []
orig_run = pl.FrameFeeder._run
def traced_run(self):
    torch.cuda.set_device(self.device_index)
    stream = torch.cuda.Stream(device=self.device_index)
    try:
        for batch in self.source:
            with torch.cuda.stream(stream):
                s = torch.cuda.Event(enable_timing=True); s.record(stream)
                t_alloc = time.perf_counter()
                d = batch.to(f'cuda:{self.device_index}', non_blocking=True)
                t_alloc = time.perf_counter() - t_alloc
                e = torch.cuda.Event(enable_timing=True); e.record(stream)
            events.append((s, e, t_alloc, time.perf_counter()))
            self.queue.put((d, e))
    finally:
        self.queue.put(None)
pl.FrameFeeder._run = traced_run
K = 30
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in pipe.run(pl.FrameFeeder((host for _ in range(K)), device=dev)): pass
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f'{K} steps, {dt / K * 1e3:.3f} ms per step')
for i, (s, e, ta, tw) in enumerate(events):
    gap = events[i - 1][1].elapsed_time(s) if i else 0.0
    print(f'batch {i:2d}: copy {s.elapsed_time(e):6.2f} ms, idle before {gap:6.2f} ms, host .to() call {ta * 1e3:6.2f} ms, issued at {(tw - t0) * 1e3:7.1f} ms')
