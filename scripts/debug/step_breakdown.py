"""What each part of the headline step costs IN the two-stream step (one GPU): full step, without
the pose parse, without detection, pose net only."""
import os, sys
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
frames = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (32, 1080, 1920, 3), dtype=np.uint8)).to(dev)
side = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]

def step(detect=True, parse=True, pose=True):
    cur = torch.cuda.current_stream(dev)
    fork = torch.cuda.Event(); fork.record(cur)
    out = []
    if pose:
        with torch.cuda.stream(side[1]):
            side[1].wait_event(fork)
            if parse:
                out.append(pose_model.estimate_async(frames).done)
            else:
                resized, _ = resize_short_side(frames, 184)
                pose_model.maps(resized)
    if detect:
        with torch.cuda.stream(side[0]):
            side[0].wait_event(fork)
            small, _ = resize_short_side(frames, 416)
            out.append(det_model.detect_async(small).done)
    cur.wait_stream(side[0]); cur.wait_stream(side[1])
    return out

def timed(tag, **kw):
    for _ in range(3): step(**kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): outs = step(**kw)
    for ev in outs: torch.cuda.current_stream(dev).wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    print(f'{tag:36s} {e0.elapsed_time(e1) / 20:7.3f} ms per step', flush=True)

timed('full step (detect + pose + parse)')
timed('without the pose parse', parse=False)
timed('without detection', detect=False)
timed('pose net + resize + export only', detect=False, parse=False)
timed('detection only', pose=False)
