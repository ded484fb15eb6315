"""Where does the end-to-end step go? (host wall clock per stage, synchronised)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('TERRAN_HOME', os.path.join(ROOT, '.pytest_cache', 'terran_home'))
os.makedirs(os.environ['TERRAN_HOME'], exist_ok=True)
import numpy as np, torch
from terran_b200 import synth
from terran_b200.face.detection import Detection
from terran_b200.face.detection.retinaface import RetinaFace
from terran_b200.face.detection.retinaface.wrapper import unpack_detections
from terran_b200.pose import Estimation
from terran_b200.pose.openpose import OpenPose
from terran_b200.pose.openpose.wrapper import unpack_poses
from terran_b200.frames import resize_short_side
from terran_b200.batching import round_faces

dev = torch.device('cuda')
det = RetinaFace(device=dev, state_dict=synth.retinaface_state_dict())
op = OpenPose(device=dev, state_dict=synth.openpose_state_dict())
host = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (32, 1080, 1920, 3), dtype=np.uint8)).pin_memory()

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r

ms, d = t(lambda: host.to(dev, non_blocking=True)); print(f'H2D 199MB            {ms:7.2f} ms  ({199.07/ms:.1f} GB/s)')
ms, small = t(lambda: resize_short_side(d, 416)[0]); print(f'resize 416           {ms:7.2f} ms')
ms, _ = t(lambda: det.forward(small)); print(f'retinaface forward   {ms:7.2f} ms')
ms, out = t(lambda: det.detect_device(small)); print(f'detect_device        {ms:7.2f} ms')
count, cand, rows = out
def d2h():
    c = count.cpu().numpy(); top = int(c.max()); return c, rows[:, :max(top, 1)].cpu().numpy()
ms, (c, r) = t(d2h); print(f'detect D2H           {ms:7.2f} ms')
ms, faces = t(lambda: unpack_detections(c, r)); print(f'unpack dicts         {ms:7.2f} ms  ({sum(len(f) for f in faces)} faces)')
ms, _ = t(lambda: round_faces(faces, 0.385)); print(f'round_faces          {ms:7.2f} ms')
ms, o = t(lambda: op.estimate_device(d)); print(f'estimate_device      {ms:7.2f} ms')
ms, _ = t(lambda: unpack_poses(*o)); print(f'unpack poses         {ms:7.2f} ms')
D = Detection(device=dev, lazy=True); D.model = det
E = Estimation(device=dev, lazy=True); E.model = op
ms, _ = t(lambda: D(d)); print(f'Detection()(cuda)    {ms:7.2f} ms')
ms, _ = t(lambda: E(d)); print(f'Estimation()(cuda)   {ms:7.2f} ms')
