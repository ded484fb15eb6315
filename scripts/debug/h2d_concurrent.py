"""Host->device rate of the 199 MB frame batch alone and concurrently with the detect+pose step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('TERRAN_HOME', os.path.join(ROOT, '.pytest_cache', 'terran_home'))
os.makedirs(os.environ['TERRAN_HOME'], exist_ok=True)
import numpy as np, torch
import bench
from terran_b200.face.detection.retinaface import RetinaFace
from terran_b200.pose.openpose import OpenPose
from terran_b200.frames import resize_short_side
dev = torch.device('cuda')
sd_det, sd_pose = bench.bench_weights()
det_model, pose_model = RetinaFace(device=dev, state_dict=sd_det), OpenPose(device=dev, state_dict=sd_pose)
host = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (32, 1080, 1920, 3), dtype=np.uint8)).pin_memory()
frames = host.to(dev)
stage = [torch.empty_like(frames) for _ in range(2)]
cs = torch.cuda.Stream(device=dev)
def copies(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record()
        for i in range(k): stage[i & 1].copy_(host, non_blocking=True)
        e1.record()
    return e0, e1
def step():
    small, _ = resize_short_side(frames, 416)
    det_model.detect_async(small); pose_model.estimate_async(frames)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = copies(20); torch.cuda.synchronize()
print(f'H2D alone: {host.numel() * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s')
e0, e1 = copies(20)
for _ in range(20): step()
torch.cuda.synchronize()
print(f'H2D during detect+pose: {host.numel() * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s')
pose_only = lambda: pose_model.net.run(resize_short_side(frames, 184)[0], 32, 184, 327, (184 * 327 * 3, 327 * 3, 3, 1))
e0, e1 = copies(20)
for _ in range(22): pose_only()
torch.cuda.synchronize()
print(f'H2D during the OpenPose net only: {host.numel() * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s')
