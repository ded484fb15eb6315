"""Debug: Detection at 32x1080p vs the oracle — where do survivors go missing?"""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2, numpy as np, torch
from oracle import detect, nets
from terran_b200 import synth
from terran_b200.face.detection.retinaface import RetinaFace
from terran_b200.frames import resize_short_side

def logit(p):
    p = np.clip(np.asarray(p, np.float64), 1e-7, 1 - 1e-7)
    return np.log(p / (1 - p))

sd = synth.retinaface_state_dict()
model = RetinaFace(device=torch.device('cuda'), state_dict=sd)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
frames = np.random.default_rng(0).integers(0, 256, (N, 1080, 1920, 3), dtype=np.uint8)
s = 416 / 1080
small = np.stack([cv2.resize(f, (int(1920 * s), int(1080 * s)), interpolation=cv2.INTER_LINEAR) for f in frames])
dev = torch.from_numpy(small).cuda()
got = [t.cpu().numpy() for t in model.heads(dev)]
want = []
for i in range(0, N, 8):
    x = torch.from_numpy(small[i:i + 8].astype(np.float32)).permute(0, 3, 1, 2).flip(1)
    want.append([t.numpy() for t in nets.retinaface_forward(sd, x)])
want = [np.concatenate([w[k] for w in want], 0) for k in range(9)]
for k in range(9):
    a, b = got[k], want[k]
    if k % 3 == 0:
        d = np.abs(logit(a) - logit(b))
        live = (np.abs(logit(a)) < 8) & (np.abs(logit(b)) < 8)
        d = np.where(live, d, 0)
    else:
        d = np.abs(a - b)
    per_frame = d.reshape(N, -1).max(1)
    print(f'head {k}: max err {d.max():.4g}; per frame', np.round(per_frame, 4).tolist())
    if k % 3 == 0:
        i = np.unravel_index(d.argmax(), d.shape)
        print('   worst at', i, 'got', a[i], 'want', b[i], 'logits', logit(a[i]), logit(b[i]))
# survivors
count, _, det = model.detect_device(dev)
count, det = count.cpu().numpy(), det.cpu().numpy()
scores, boxes, lmks = detect.decode(want, *small.shape[1:3])
gs, gb, gl = detect.decode(got, *small.shape[1:3])
ref = detect.select(scores, boxes, lmks)
mine = detect.select(gs, gb, gl)
for n in range(min(N, 12)):
    ours = set(det[n, :count[n], 15].view(np.int32).tolist())
    r = set(ref[n]['index'].tolist())
    m = set(mine[n]['index'].tolist())
    print(f'frame {n}: ref {len(r)} ours {len(ours)} oracle-on-our-heads {len(m)} | ours==oracle(ourheads) {ours == m} | ref-ours {sorted(r - ours)[:8]} ours-ref {sorted(ours - r)[:8]}')
    for i in sorted(r - ours)[:4]:
        print(f'    missing {i}: ref score {scores[n, i]:.4f} (logit {logit(scores[n, i]):.3f}) our score {gs[n, i]:.4f} (logit {logit(gs[n, i]):.3f}) box ref {boxes[n, i]} ours {gb[n, i]}')

def iou(a, b):
    iw = max(0.0, min(a[2], b[2]) - max(a[0], b[0])); ih = max(0.0, min(a[3], b[3]) - max(a[1], b[1]))
    inter = iw * ih
    return inter / ((a[2]-a[0])*(a[3]-a[1]) + (b[2]-b[0])*(b[3]-b[1]) - inter)

print('---- cascade analysis')
for n in range(N):
    ours = set(det[n, :count[n], 15].view(np.int32).tolist())
    r = set(ref[n]['index'].tolist())
    for j in sorted(ours - r):
        print(f'frame {n}: extra {j}: ref score {scores[n, j]:.4f} ours {gs[n, j]:.4f}')
        for k in sorted(r):
            if scores[n, k] > scores[n, j]:
                a, b = iou(boxes[n, k], boxes[n, j]), iou(gb[n, k], gb[n, j])
                if a > 0.3 or b > 0.3:
                    print(f'     vs ref survivor {k} (score {scores[n, k]:.4f}/{gs[n, k]:.4f}): IoU ref {a:.5f} ours {b:.5f}')
        # candidates (not survivors) in ref with higher score overlapping j
        cand = np.flatnonzero(scores[n] >= 0.5)
        for k in cand:
            if k not in r and scores[n, k] > scores[n, j]:
                a, b = iou(boxes[n, k], boxes[n, j]), iou(gb[n, k], gb[n, j])
                if a > 0.38:
                    print(f'     vs ref NON-survivor cand {k} (score {scores[n, k]:.4f}/{gs[n, k]:.4f}): IoU ref {a:.5f} ours {b:.5f}')
