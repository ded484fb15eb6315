"""Why does the end-to-end window slow down when many ranks share one host?

Run under torchrun.  Prints the host's CPU / cgroup / NUMA picture, then repeats bench.py's
e2e window (K steps through FrameFeeder + PerceptionPipeline) several times per rank in two
modes — spinning event waits (the CUDA default) and blocking ones — and prints every rank's
window times, its per-step host timestamps for the slowest window and the process CPU time
burnt per window.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import bench  # noqa: E402


def host_picture():
    out = {'cpu_count': os.cpu_count(), 'affinity': len(os.sched_getaffinity(0))}
    for path in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us',
                 '/sys/fs/cgroup/cpu/cpu.cfs_period_us', '/sys/fs/cgroup/cpuset.cpus.effective',
                 '/sys/devices/system/node/online', '/proc/loadavg'):
        try:
            with open(path) as f:
                out[path] = f.read().strip()
        except OSError:
            pass
    try:
        import subprocess
        out['topo'] = subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True,
                                     timeout=20).stdout
    except Exception as e:          # noqa: BLE001
        out['topo'] = repr(e)
    try:
        import pynvml
        pynvml.nvmlInit()
        aff = []
        for i in range(pynvml.nvmlDeviceGetCount()):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if word >> b & 1]
            aff.append((i, len(cpus), cpus[:1], cpus[-1:]))
        out['nvml_cpu_affinity'] = aff
    except Exception as e:          # noqa: BLE001
        out['nvml_cpu_affinity'] = repr(e)
    return out


def main():
    from terran_b200 import defaults, parallel
    from terran_b200.face.detection import Detection
    from terran_b200.face.detection.retinaface import RetinaFace
    from terran_b200.pipeline import FrameFeeder, PerceptionPipeline
    from terran_b200.pose import Estimation
    from terran_b200.pose.openpose import OpenPose
    import torch.distributed as dist

    steps = int(os.environ.get('STEPS', '10'))
    windows = int(os.environ.get('WINDOWS', '4'))
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    binding = parallel.bind_to_gpu_numa(local, world)
    if rank == 0:
        for k, v in host_picture().items():
            print(f'[host] {k}: {v}', flush=True)
    w_det, w_pose = bench.bench_weights() if rank == 0 else (None, None)
    det_model = RetinaFace(device=dev, state_dict=parallel.broadcast_state_dict(w_det))
    pose_model = OpenPose(device=dev, state_dict=parallel.broadcast_state_dict(w_pose))
    detection = Detection(device=dev, lazy=True)
    detection.model = det_model
    estimation = Estimation(device=dev, lazy=True)
    estimation.model = pose_model
    H, W = bench.FRAME_HW
    host = torch.from_numpy(np.random.default_rng(rank).integers(
        0, 256, (bench.BATCH, H, W, 3), dtype=np.uint8)).pin_memory()
    pipe = PerceptionPipeline(detection, estimation, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    report = {'rank': rank, 'binding': binding}
    modes = os.environ.get('MODES', 'spin,blocking,spin').split(',')
    for mode in modes:
        defaults.blocking_events = mode == 'blocking'
        os.environ['TRB_FEEDER_PRIVATE_STREAM'] = '1' if mode == 'private' else '0'
        for _ in pipe.run(FrameFeeder((host for _ in range(8)), device=dev)):
            pass
        rows = []
        for _ in range(windows):
            barrier()
            c0 = time.process_time()
            t0 = time.perf_counter()
            stamps = []
            for _r in pipe.run(FrameFeeder((host for _ in range(steps)), device=dev)):
                stamps.append(round((time.perf_counter() - t0) * 1e3, 1))
            torch.cuda.synchronize()
            rows.append((round((time.perf_counter() - t0) * 1e3, 1),
                         round((time.process_time() - c0) * 1e3, 1), stamps))
        report.setdefault(mode, []).append(rows)
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, report)
    else:
        gathered = [report]
    if rank == 0:
        for mode in dict.fromkeys(modes):
            for rep, _ in enumerate(gathered[0][mode]):
                print(f'== {mode} (pass {rep}): per rank, window ms (process CPU ms)')
                worst = (0.0, None)
                for g in gathered:
                    rows = g[mode][rep]
                    print(f"  rank {g['rank']} {g['binding']}: "
                          + ' '.join(f'{t}({c})' for t, c, _ in rows))
                    for t, _, stamps in rows:
                        if t > worst[0]:
                            worst = (t, (g['rank'], stamps))
                per_window = [max(g[mode][rep][w][0] for g in gathered) for w in range(windows)]
                fps = [round(world * bench.BATCH * steps / (t / 1e3)) for t in per_window]
                print(f'  max over ranks per window: {per_window} -> frames/s {fps}')
                print(f'  slowest window: rank {worst[1][0]} step completion stamps {worst[1][1]}')
                print('  first result of each window, rank 0 (ms):', [r[2][0] for r in gathered[0][mode][rep]])
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
