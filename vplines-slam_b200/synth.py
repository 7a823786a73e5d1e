"""Synthetic workloads of BASELINE.json: deterministic frame
sequences shaped like the reference's sensors (EuRoC 752x480, D455-like 1280x720,
1920x1080 Manhattan scenes).  numpy only, seeded.  Used by bench.py and the tests;
not part of the compute path.

Scene = smooth background (low-pass noise, +-30 around 128) + filled convex quads +
anti-aliased strokes, rendered once on a canvas larger than the frame; each frame is
a view of the canvas under a slowly drifting similarity transform (translation
U[-3,3] px, rotation U[-0.3,0.3] deg, scale U[0.998,1.002] per frame, cumulative) with
bilinear sampling, plus N(0, 2^2) pixel noise.
"""
import numpy as np


def _smooth_noise(rng, h, w, cell=64, amp=30.0):
    gh, gw = h // cell + 3, w // cell + 3
    g = rng.uniform(-1, 1, (gh, gw))
    ys = np.arange(h) / cell
    xs = np.arange(w) / cell
    y0 = ys.astype(int); x0 = xs.astype(int)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    fy = fy * fy * (3 - 2 * fy); fx = fx * fx * (3 - 2 * fx)
    a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
    return amp * ((a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy)


def _draw_convex(canvas, pts, value):
    """Anti-aliased fill of a convex polygon (vertices in order) by edge coverage."""
    h, w = canvas.shape
    pts = np.asarray(pts, np.float64)
    area = 0.5 * np.sum(pts[:, 0] * np.roll(pts[:, 1], -1) - np.roll(pts[:, 0], -1) * pts[:, 1])
    if abs(area) < 1e-6:
        return
    if area < 0:
        pts = pts[::-1]
    x0 = int(max(0, np.floor(pts[:, 0].min()) - 1)); x1 = int(min(w, np.ceil(pts[:, 0].max()) + 2))
    y0 = int(max(0, np.floor(pts[:, 1].min()) - 1)); y1 = int(min(h, np.ceil(pts[:, 1].max()) + 2))
    if x1 <= x0 or y1 <= y0:
        return
    yy, xx = np.mgrid[y0:y1, x0:x1]
    cov = np.ones(yy.shape)
    n = len(pts)
    for i in range(n):
        p, q = pts[i], pts[(i + 1) % n]
        e = q - p
        ln = np.hypot(e[0], e[1])
        if ln < 1e-9:
            continue
        d = ((xx - p[0]) * e[1] - (yy - p[1]) * e[0]) / ln  # >0 outside for CCW in image coords
        cov = np.minimum(cov, np.clip(0.5 - d, 0, 1))
    sub = canvas[y0:y1, x0:x1]
    canvas[y0:y1, x0:x1] = sub * (1 - cov) + value * cov


def _stroke(p, ang, length, thick):
    d = np.array([np.cos(ang), np.sin(ang)]); n = np.array([-d[1], d[0]])
    a = p - d * length / 2; b = p + d * length / 2
    t = thick / 2
    return [a + n * t, b + n * t, b - n * t, a - n * t]


def make_canvas(w, h, seed, n_quads=40, n_strokes=60, manhattan=False, margin=96):
    rng = np.random.default_rng(seed)
    W, H = w + 2 * margin, h + 2 * margin
    canvas = 128.0 + _smooth_noise(rng, H, W)
    if manhattan:
        # three families of segments converging to three vanishing points
        vps = [np.array([W / 2 + rng.uniform(-0.3, 0.3) * W, -3.0 * H]),
               np.array([-2.5 * W, H / 2 + rng.uniform(-0.3, 0.3) * H]),
               np.array([3.5 * W, H / 2 + rng.uniform(-0.3, 0.3) * H])]
        for i in range(n_strokes):
            vp = vps[i % 3]
            p = np.array([rng.uniform(0, W), rng.uniform(0, H)])
            dirv = vp - p
            ang = np.arctan2(dirv[1], dirv[0])
            _draw_convex(canvas, _stroke(p, ang, rng.uniform(25, 400), rng.integers(2, 7)), rng.uniform(20, 235))
        return canvas
    for _ in range(n_quads):
        c = np.array([rng.uniform(0, W), rng.uniform(0, H)])
        r = rng.uniform(30, 220)
        angs = np.sort(rng.uniform(0, 2 * np.pi, 4))
        pts = [c + r * rng.uniform(0.5, 1.0) * np.array([np.cos(a), np.sin(a)]) for a in angs]
        _draw_convex(canvas, pts, rng.uniform(20, 235))
    for _ in range(n_strokes):
        p = np.array([rng.uniform(0, W), rng.uniform(0, H)])
        _draw_convex(canvas, _stroke(p, rng.uniform(0, np.pi), rng.uniform(30, 300), rng.integers(1, 5)),
                     rng.uniform(20, 235))
    return canvas


def _sample(canvas, A, w, h):
    """Bilinear sample canvas at A @ [x, y, 1] for the w x h frame grid."""
    H, W = canvas.shape
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    sx = A[0, 0] * xs + A[0, 1] * ys + A[0, 2]
    sy = A[1, 0] * xs + A[1, 1] * ys + A[1, 2]
    sx = np.clip(sx, 0, W - 1.001); sy = np.clip(sy, 0, H - 1.001)
    x0 = sx.astype(np.int64); y0 = sy.astype(np.int64)
    fx = sx - x0; fy = sy - y0
    a = canvas[y0, x0]; b = canvas[y0, x0 + 1]; c = canvas[y0 + 1, x0]; d = canvas[y0 + 1, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def sequence(n_frames, w=752, h=480, seed=20240601, manhattan=False, n_quads=40, n_strokes=60, margin=96):
    """(n_frames, h, w) uint8."""
    canvas = make_canvas(w, h, seed, n_quads=n_quads, n_strokes=n_strokes, manhattan=manhattan, margin=margin)
    rng = np.random.default_rng(seed + 7)
    out = np.empty((n_frames, h, w), np.uint8)
    tx, ty, rot, sc = float(margin), float(margin), 0.0, 1.0
    cx, cy = w / 2.0, h / 2.0
    for f in range(n_frames):
        if f > 0:
            tx += rng.uniform(-3, 3); ty += rng.uniform(-3, 3)
            rot += np.deg2rad(rng.uniform(-0.3, 0.3)); sc *= rng.uniform(0.998, 1.002)
            # keep the view inside the canvas
            tx = float(np.clip(tx, margin * 0.35, margin * 1.65)); ty = float(np.clip(ty, margin * 0.35, margin * 1.65))
            rot = float(np.clip(rot, -0.05, 0.05)); sc = float(np.clip(sc, 0.97, 1.03))
        c, s = np.cos(rot) * sc, np.sin(rot) * sc
        A = np.array([[c, -s, tx + cx - (c * cx - s * cy)], [s, c, ty + cy - (s * cx + c * cy)]])
        img = _sample(canvas, A, w, h)
        noise = np.random.default_rng(seed + 1000 + f).normal(0.0, 2.0, (h, w))
        out[f] = np.clip(np.rint(img + noise), 0, 255).astype(np.uint8)
    return out


CONFIGS = {
    # name: (w, h, octaves, k, generator kwargs)
    "C2_euroc_752x480": dict(w=752, h=480, num_octaves=1, k=1, gen=dict(n_quads=40, n_strokes=60)),
    "C3_d455_1280x720": dict(w=1280, h=720, num_octaves=2, k=2, gen=dict(n_quads=70, n_strokes=110)),
    "C4_manhattan_1920x1080": dict(w=1920, h=1080, num_octaves=1, k=2, gen=dict(manhattan=True, n_strokes=560)),
}


def config_sequence(name, n_frames, seed=20240601):
    c = CONFIGS[name]
    return sequence(n_frames, w=c["w"], h=c["h"], seed=seed, **c["gen"])
