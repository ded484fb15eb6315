"""Where does the detect+recog+pose pipeline spend its time (one GPU)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('TERRAN_HOME', os.path.join(ROOT, '.pytest_cache', 'terran_home'))
os.makedirs(os.environ['TERRAN_HOME'], exist_ok=True)
import numpy as np, torch
from terran_b200 import synth
from terran_b200.face.recognition.arcface import ArcFace
dev = torch.device('cuda')
arc = ArcFace(device=dev, state_dict=synth.arcface_state_dict())
for n in (256, 512, 864, 861):
    crops = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (n, 3, 112, 112), dtype=np.uint8)).to(dev)
    for _ in range(2): arc.embed_device(crops, 'nchw_bgr')
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): arc.embed_device(crops, 'nchw_bgr')
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f'arcface N={n}: {dt*1e3:.2f} ms  {n/dt:.0f} crops/s', flush=True)
import bench
from terran_b200.face.detection.retinaface import RetinaFace
from terran_b200.pose.openpose import OpenPose
sd_det, sd_pose = bench.bench_weights()
det_model, pose_model = RetinaFace(device=dev, state_dict=sd_det), OpenPose(device=dev, state_dict=sd_pose)
frames = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (32, 1080, 1920, 3), dtype=np.uint8)).to(dev)
for k in range(2):
    print(bench.pipeline_with_recognition(det_model, pose_model, arc, frames, dev), flush=True)
