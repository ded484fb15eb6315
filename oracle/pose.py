"""ORACLE (test infrastructure, not product code): numpy restatement of the
OpenPose heat-map / PAF parse — x8 bicubic upsample, peak extraction, limb
line-integral scoring, greedy matching, human assembly and keypoint output.

Follows ``terran/pose/openpose/wrapper.py:182-485`` (and ``get_keypoints``
``:37-90``, ``build_segments`` ``:125-163``) of the reference, restated in the
kernel-shaped loop form of SURVEY.md Appendix A.3/A.4/A.8.  The fp32 operation
ORDER written here (no fused multiply-add, sequential sums) is the contract the
CUDA kernels reproduce bit-for-bit; the reference's own torch reductions may
differ from it in the last ulp of a score, never in an index on the committed
golden scenes (``tests/golden/openpose_parse_*.npz``, produced by running the
reference's ``OpenPose.call`` through ``oracle/make_golden.py``).
"""
import numpy as np

F32 = np.float32

MAP_IDX = (
    (31, 32), (39, 40), (33, 34), (35, 36), (41, 42), (43, 44),
    (19, 20), (21, 22), (23, 24), (25, 26), (27, 28), (29, 30),
    (47, 48), (49, 50), (53, 54), (51, 52), (55, 56), (37, 38),
    (45, 46),
)
LIMBSEQ = (
    (2, 3), (2, 6), (3, 4), (4, 5), (6, 7), (7, 8), (2, 9),
    (9, 10), (10, 11), (2, 12), (12, 13), (13, 14), (2, 1),
    (1, 15), (15, 17), (1, 16), (16, 18), (3, 17), (6, 18),
)
KEYPOINT_THRESHOLD = F32(0.1)
MIDPOINT_THRESHOLD = F32(0.05)
HUMAN_THRESHOLD = 0.4
NUM_MIDPOINTS = 10
UP = 8


def bicubic_table():
    """(8,4) f32 weights of taps f-1..f+2 for output phase o%8 (A=-0.75,
    align_corners=False): src = (o+0.5)/8-0.5, t = src-floor(src) is an exact
    multiple of 1/16, so 8 phases cover every output pixel."""
    A = F32(-0.75)
    tab = np.zeros((UP, 4), F32)

    def c1(v):     # |v| <= 1
        return ((A + F32(2)) * v - (A + F32(3))) * v * v + F32(1)

    def c2(v):     # 1 < |v| < 2
        return ((A * v - F32(5) * A) * v + F32(8) * A) * v - F32(4) * A

    for ph in range(UP):
        src = F32((ph + 0.5) / UP - 0.5)
        t = F32(src - np.floor(src))
        tab[ph] = (c2(t + F32(1)), c1(t), c1(F32(1) - t), c2(F32(2) - t))
    return tab


def _taps(n_in):
    """For every output index: the 4 clamped source indices and the phase."""
    o = np.arange(n_in * UP)
    f = np.floor((o + 0.5) / UP - 0.5).astype(np.int64)
    idx = np.clip(f[:, None] + np.arange(-1, 3)[None, :], 0, n_in - 1)
    return idx, o % UP


def bicubic_up8(maps):
    """(C,h,w) f32 -> (C,8h,8w) f32.  value = sum_i wy_i * (sum_j wx_j * s_ij),
    both sums left to right, every product and sum rounded to f32."""
    maps = np.asarray(maps, F32)
    tab = bicubic_table()
    C, h, w = maps.shape
    iy, py = _taps(h)
    ix, px = _taps(w)
    wx = tab[px]                                  # (W,4)
    wy = tab[py]                                  # (H,4)
    g = maps[:, :, ix]                            # (C,h,W,4)
    rows = g[..., 0] * wx[:, 0]
    for j in range(1, 4):
        rows = rows + g[..., j] * wx[:, j]        # (C,h,W) f32
    r = rows[:, iy, :]                            # (C,H,4,W)
    out = r[:, :, 0, :] * wy[None, :, 0, None]
    for i in range(1, 4):
        out = out + r[:, :, i, :] * wy[None, :, i, None]
    return out.astype(F32)


def find_peaks(heat_up):
    """heat_up (>=18,H,W) f32.  Returns per part (locs (n,2) int64 (y,x) in
    row-major order, scores (n,) f32) — wrapper.py:235-262."""
    locs, scores = [], []
    for part in range(18):
        m = heat_up[part]
        c = m[1:-1, 1:-1]
        mask = ((c >= m[:-2, 1:-1]) & (c >= m[1:-1, :-2]) & (c >= m[2:, 1:-1])
                & (c >= m[1:-1, 2:]) & (c >= KEYPOINT_THRESHOLD))
        p = np.argwhere(mask) + 1
        locs.append(p.astype(np.int64))
        scores.append(m[p[:, 0], p[:, 1]].astype(F32))
    return locs, scores


def segment_points(a, b):
    """torch.linspace(a, b, 10) in f32 truncated to int (Appendix A.4)."""
    a, b = F32(a), F32(b)
    step = F32((b - a) / F32(NUM_MIDPOINTS - 1))
    pts = []
    for k in range(NUM_MIDPOINTS):
        if k < NUM_MIDPOINTS // 2:
            v = F32(a + F32(step * F32(k)))
        else:
            v = F32(b - F32(step * F32(NUM_MIDPOINTS - 1 - k)))
        pts.append(int(v))
    return pts


def limb_candidates(paf_up, limb, loc_src, loc_dst):
    """All accepted (i, j, reg) for one limb, row-major (i, j) order —
    wrapper.py:274-333."""
    cx, cy = MAP_IDX[limb][0] - 19, MAP_IDX[limb][1] - 19
    H_up = paf_up.shape[1]
    cand = []
    for i, (sy, sx) in enumerate(loc_src):
        for j, (dy, dx) in enumerate(loc_dst):
            vy, vx = F32(dy - sy), F32(dx - sx)
            n = F32(np.sqrt(F32(F32(vy * vy) + F32(vx * vx))))
            with np.errstate(divide='ignore', invalid='ignore'):
                uy, ux = F32(vy / n), F32(vx / n)
                ys = segment_points(sy, dy)
                xs = segment_points(sx, dx)
                total = F32(0)
                above = 0
                for k in range(NUM_MIDPOINTS):
                    m = F32(F32(paf_up[cx, ys[k], xs[k]] * ux)
                            + F32(paf_up[cy, ys[k], xs[k]] * uy))
                    if m > MIDPOINT_THRESHOLD:
                        above += 1
                    total = F32(total + m)
                pen = F32(min(F32(F32(F32(0.5 * H_up) / n) - F32(1)), F32(0)))
                reg = F32(F32(total / F32(NUM_MIDPOINTS)) + pen)
            if above > 0.8 * NUM_MIDPOINTS and reg > 0:
                cand.append((i, j, reg))
    return cand


def greedy_match(cand, ids_src, ids_dst):
    """wrapper.py:335-366: descending by score (stable; inputs tie-free), ONE
    shared ``seen`` set for source and destination indices, early ``break``
    before the indices of the last accepted pair are inserted."""
    order = sorted(range(len(cand)), key=lambda t: -float(cand[t][2]))
    limit = min(len(ids_src), len(ids_dst))
    seen = set()
    conns = []
    for t in order:
        i, j, s = cand[t]
        if i not in seen and j not in seen:
            conns.append((float(ids_src[i]), float(ids_dst[j]), float(s)))
            if len(conns) >= limit:
                break
            seen.add(i)
            seen.add(j)
    return np.asarray(conns, np.float64).reshape(-1, 3)


def assemble(all_conns, missing, peak_scores_by_id):
    """wrapper.py:380-478: rows of 20 f64 (18 peak ids or -1, score sum, count)."""
    humans = np.zeros((0, 20))
    for limb in range(19):
        if limb in missing:
            continue
        conns = all_conns[limb]
        ks, kd = LIMBSEQ[limb][0] - 1, LIMBSEQ[limb][1] - 1
        for c in range(len(conns)):
            ps, pd, sc = conns[c]
            matched = [h for h in range(len(humans))
                       if humans[h, ks] == ps or humans[h, kd] == pd]
            if len(matched) == 1:
                h = humans[matched[0]]
                if h[kd] != pd:
                    h[kd] = pd
                    h[-1] += 1
                    h[-2] += peak_scores_by_id[int(pd)] + sc
            elif len(matched) == 2:
                h1, h2 = humans[matched[0]], humans[matched[1]]
                both = ((h1[:-2] >= 0).astype(int) + (h2[:-2] >= 0).astype(int)) == 2
                if not both.any():
                    h1[:-2] += h2[:-2] + 1
                    h1[-2:] += h2[-2:]
                    h1[-2] += sc
                    humans = np.delete(humans, matched[1], 0)
                else:
                    h1[kd] = pd
                    h1[-1] += 1
                    h1[-2] += peak_scores_by_id[int(pd)] + sc
            elif not matched and limb < 17:
                row = -np.ones(20)
                row[ks], row[kd] = ps, pd
                row[-1] = 2
                row[-2] = (0 + peak_scores_by_id[int(ps)] + peak_scores_by_id[int(pd)]) + sc
                humans = np.vstack([humans, row])
    keep = [h for h in range(len(humans))
            if not (humans[h, -1] < 4 or humans[h, -2] / humans[h, -1] < HUMAN_THRESHOLD)]
    return humans[keep]


def parse_frame(paf, heat, scale):
    """paf (38,h,w), heat (19,h,w) f32 network outputs of ONE frame.  Returns
    the reference's list of {'keypoints': int32 (18,3), 'score': f64}."""
    heat_up = bicubic_up8(heat)
    paf_up = bicubic_up8(paf)
    locs, scores = find_peaks(heat_up)
    ids, n = [], 0
    for p in locs:
        ids.append(np.arange(n, n + len(p)))
        n += len(p)
    all_conns, missing = [], []
    for limb in range(19):
        ks, kd = LIMBSEQ[limb][0] - 1, LIMBSEQ[limb][1] - 1
        if len(locs[ks]) == 0 or len(locs[kd]) == 0:
            missing.append(limb)
            all_conns.append(np.zeros((0, 3)))
            continue
        cand = limb_candidates(paf_up, limb, locs[ks], locs[kd])
        all_conns.append(greedy_match(cand, ids[ks], ids[kd]))
    flat_locs = np.concatenate(locs, 0) if n else np.zeros((0, 2), np.int64)
    flat_scores = (np.concatenate(scores).astype(np.float64) if n
                   else np.zeros(0, np.float64))
    humans = assemble(all_conns, missing, flat_scores)
    out = []
    for h in humans:
        kp = np.zeros((18, 3), np.int32)
        for j in range(18):
            pid = int(np.int32(h[j]))
            if pid != -1:
                y, x = flat_locs[pid].astype(np.float64)
                kp[j] = (np.int32(x / scale), np.int32(y / scale), 1)
        out.append({'keypoints': kp, 'score': np.float64(h[-2] / h[-1])})
    return out


def parse(pafs, heats, scale):
    return [parse_frame(p, h, scale) for p, h in zip(np.asarray(pafs, F32),
                                                     np.asarray(heats, F32))]


# ------------------------------------------------------------ synthetic scenes

#: Rough skeleton template in a unit box: (x, y) for the 18 COCO joints.
_TEMPLATE = np.array([
    (0.50, 0.08), (0.50, 0.22), (0.36, 0.22), (0.30, 0.40), (0.27, 0.56),
    (0.64, 0.22), (0.70, 0.40), (0.73, 0.56), (0.42, 0.55), (0.41, 0.75),
    (0.40, 0.95), (0.58, 0.55), (0.59, 0.75), (0.60, 0.95), (0.46, 0.05),
    (0.54, 0.05), (0.41, 0.08), (0.59, 0.08),
])


def synthetic_scene(seed, h=23, w=40, people=None, drop=0.1):
    """Seeded (paf (38,h,w), heat (19,h,w)) f32 with jittered skeletons:
    Gaussian joint blobs and unit-vector PAF ribbons."""
    rng = np.random.default_rng(seed)
    P = int(rng.integers(1, 9)) if people is None else people
    heat = np.zeros((19, h, w), np.float64)
    paf = np.zeros((38, h, w), np.float64)
    cnt = np.zeros((38, h, w), np.float64)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    for _ in range(P):
        size = rng.uniform(0.45, 0.95) * h
        ox = rng.uniform(-0.1 * size, w - 0.6 * size)
        oy = rng.uniform(-0.05 * h, h - 0.9 * size)
        pts = _TEMPLATE * [size * 0.62, size] + [ox, oy]
        pts = pts + rng.normal(0, 0.25, pts.shape)
        present = rng.random(18) >= drop
        amp = rng.uniform(0.7, 1.0, 18)
        for j in range(18):
            if present[j]:
                x, y = pts[j]
                heat[j] = np.maximum(
                    heat[j], amp[j] * np.exp(-((xx - x) ** 2 + (yy - y) ** 2) / (2 * 0.8 ** 2)))
        for l in range(19):
            a, b = LIMBSEQ[l][0] - 1, LIMBSEQ[l][1] - 1
            if not (present[a] and present[b]):
                continue
            pa, pb = pts[a], pts[b]
            v = pb - pa
            ln = np.linalg.norm(v)
            if ln < 1e-6:
                continue
            u = v / ln
            rx, ry = xx - pa[0], yy - pa[1]
            along = rx * u[0] + ry * u[1]
            perp = np.abs(rx * u[1] - ry * u[0])
            mask = (along >= -0.5) & (along <= ln + 0.5) & (perp <= 1.0)
            cxi, cyi = MAP_IDX[l][0] - 19, MAP_IDX[l][1] - 19
            strength = rng.uniform(0.75, 1.0)
            paf[cxi][mask] += u[0] * strength
            paf[cyi][mask] += u[1] * strength
            cnt[cxi][mask] += 1
            cnt[cyi][mask] += 1
    paf = paf / np.maximum(cnt, 1)
    heat += rng.uniform(0, 0.02, heat.shape)      # tie-breaking noise, heat >= 0
    paf += rng.normal(0, 0.01, paf.shape)
    # fp16-representable so fixtures can be stored losslessly at half size.
    return (paf.astype(np.float16).astype(F32), heat.astype(np.float16).astype(F32))
