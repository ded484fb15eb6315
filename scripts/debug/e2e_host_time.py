"""Host time per e2e step: submit (enqueue) vs result (wait + unpack), one GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('TERRAN_HOME', os.path.join(ROOT, '.pytest_cache', 'terran_home'))
os.makedirs(os.environ['TERRAN_HOME'], exist_ok=True)
import numpy as np, torch
import bench
from terran_b200.face.detection import Detection
from terran_b200.face.detection.retinaface import RetinaFace
from terran_b200.pose import Estimation
from terran_b200.pose.openpose import OpenPose
from terran_b200.pipeline import FrameFeeder, PerceptionPipeline
dev = torch.device('cuda')
sd_det, sd_pose = bench.bench_weights()
det = Detection(device=dev, lazy=True); det.model = RetinaFace(device=dev, state_dict=sd_det)
est = Estimation(device=dev, lazy=True); est.model = OpenPose(device=dev, state_dict=sd_pose)
host = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (32, 1080, 1920, 3), dtype=np.uint8)).pin_memory()
pipe = PerceptionPipeline(det, est, device=dev)
for _ in pipe.run(FrameFeeder((host for _ in range(3)), device=dev)): pass
K = 40
t_sub = t_res = t_get = 0.0
torch.cuda.synchronize(); t0 = time.perf_counter()
prev = None
it = iter(FrameFeeder((host for _ in range(K)), device=dev))
while True:
    a = time.perf_counter()
    try: frames = next(it)
    except StopIteration: break
    b = time.perf_counter(); t_get += b - a
    cur = pipe.submit(frames)
    c = time.perf_counter(); t_sub += c - b
    if prev is not None:
        prev.result(); t_res += time.perf_counter() - c
    prev = cur
prev.result()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f'{K} steps: {dt/K*1e3:.3f} ms per step; host: feeder wait {t_get/K*1e3:.3f}, submit {t_sub/K*1e3:.3f}, result (wait + unpack) {t_res/K*1e3:.3f} ms per step')
# unpack cost alone
r = cur.handles
t1 = time.perf_counter()
for _ in range(10):
    from terran_b200.face.detection.retinaface.wrapper import unpack_detections
    p = cur.handles[0]
t2 = time.perf_counter()
