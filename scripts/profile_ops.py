"""Per-op device timing of the three nets at the BASELINE shapes (CUDA events
around every launch, tr_net_profile).  Usage on the GPU box:
    python scripts/profile_ops.py [retinaface|openpose|arcface] [--reps 5]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('TERRAN_HOME', os.path.join(ROOT, '.pytest_cache', 'terran_home'))
os.makedirs(os.environ['TERRAN_HOME'], exist_ok=True)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from terran_b200 import _native as nat, synth  # noqa: E402

NAMES = {0: 'stem', 1: 'conv', 2: 'dw', 3: 'pool', 4: 'copy', 5: 'view', 6: 'sep'}


def profile(model_name, reps):
    dev = torch.device('cuda')
    rng = np.random.default_rng(0)
    if model_name == 'retinaface':
        from terran_b200.face.detection.retinaface import RetinaFace
        m = RetinaFace(device=dev, state_dict=synth.retinaface_state_dict())
        x = torch.from_numpy(rng.integers(0, 256, (32, 416, 739, 3), dtype=np.uint8)).to(dev)
        run = lambda: m.forward(x)
        net = m.net
    elif model_name == 'openpose':
        from terran_b200.pose.openpose import OpenPose
        m = OpenPose(device=dev, state_dict=synth.openpose_state_dict())
        x = torch.from_numpy(rng.integers(0, 256, (32, 184, 327, 3), dtype=np.uint8)).to(dev)
        N, H, W, _ = x.shape
        run = lambda: m.net.run(x, N, H, W, (H * W * 3, W * 3, 3, 1))
        net = m.net
    else:
        from terran_b200.face.recognition.arcface import ArcFace
        m = ArcFace(device=dev, state_dict=synth.arcface_state_dict())
        x = torch.from_numpy(rng.integers(0, 256, (256, 112, 112, 3), dtype=np.uint8)).to(dev)
        run = lambda: m.embed_device(x)
        net = m.net
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    whole = e0.elapsed_time(e1) / 10
    net.set_profile(True)
    acc = None
    for _ in range(reps):
        run()
        prof = net.profile()
        ms = np.array([p[0] for p in prof])
        acc = ms if acc is None else acc + ms
    net.set_profile(False)
    acc /= reps
    ops = [op for op in net.program.ops if op.type != nat.TR_OP_VIEW]
    total = acc.sum()
    print(f'== {model_name}: {len(ops)} launches, {total:.3f} ms per batch of {x.shape[0]} summed over '
          f'per-op events ({x.shape[0] / total * 1e3:.0f} units/s); {whole:.3f} ms back to back '
          f'({x.shape[0] / whole * 1e3:.0f} units/s)')
    tc_ms = tc_fl = 0.0
    groups = {}
    for i, (op, (_, is_tc, flops)) in enumerate(zip(ops, prof)):
        _, n, h, w, c = net.buffer_info(op.out)
        tf = flops / (acc[i] * 1e-3) / 1e12 if acc[i] > 0 else 0
        if is_tc:
            tc_ms += acc[i]
            tc_fl += flops
        key = (NAMES[op.type], bool(is_tc), op.k, op.stride, op.cin_real or op.in_c,
               op.cout_real or op.out_c, h, w, op.res >= 0, op.out2 >= 0)
        g = groups.setdefault(key, [0, 0.0, 0.0])
        g[0] += 1; g[1] += acc[i]; g[2] += flops
        if BRIEF:
            continue
        print(f'{i:3d} {NAMES[op.type]:5s}{"*" if is_tc else " "} k{op.k} s{op.stride} '
              f'{op.cin_real or op.in_c:4d}->{op.cout_real or op.out_c:4d} out {h:3d}x{w:3d} '
              f'{acc[i] * 1e3:8.1f} us {100 * acc[i] / total:5.1f}% {tf:7.1f} TFLOP/s')
    if BRIEF:
        for key, (cnt, ms, fl) in groups.items():
            name, is_tc, k, st, ci, co, h, w, has_res, has_out2 = key
            print(f'{cnt:3d}x {name:5s}{"*" if is_tc else " "} k{k} s{st} {ci:4d}->{co:4d} out {h:3d}x{w:3d}'
                  f'{" +res" if has_res else ""}{" +out2" if has_out2 else ""}: {ms / cnt * 1e3:8.1f} us each, '
                  f'{100 * ms / total:5.1f}%, {fl / (ms * 1e-3) / 1e12 if ms else 0:7.1f} TFLOP/s')
    if tc_ms:
        print(f'   tcgen05 total: {tc_ms:.3f} ms, {tc_fl / 1e9:.1f} GFLOP, '
              f'{tc_fl / (tc_ms * 1e-3) / 1e12:.1f} TFLOP/s')


BRIEF = '--brief' in sys.argv

if __name__ == '__main__':
    names = [a for a in sys.argv[1:] if not a.startswith('--')] or ['retinaface', 'openpose']
    reps = 5
    for n in names:
        profile(n, reps)
