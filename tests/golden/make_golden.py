"""Generates the golden vectors under tests/golden/ (run once, in the authoring
container, where /root/reference and cv2 4.13 exist; the GPU box has neither need):

  mh04_frames.npz   the 15 bundled EuRoC MH_04 frames of the reference
                    (/root/reference/line_matching/data/mh04/imgs/1..15.png, 752x480 u8) --
                    input data for config C1, not source code.
  cv2_lsd.npz       cv2 4.13 createLineSegmentDetector(LSD_REFINE_ADV).detect on
                    GaussianBlur(5x5, sigma 1) of frames 1, 5, 10: lines / width / prec / nfa,
                    plus REFINE_STD and REFINE_NONE segment arrays for frame 1 and the
                    known-answer step-edge cases of SURVEY.md section 8c.
  cv2_prims.npz     cv2 4.13 GaussianBlur(5x5,1), GaussianBlur(7x7,.75), resize(.8,
                    INTER_LINEAR_EXACT), pyrDown(w/2,h/2), Sobel dx/dy on a seeded random
                    image (full arrays) and sha256 of the same on frame 1; fastAtan2 samples.
  cv2_preproc.npz   cv2 4.13 remap(INTER_LINEAR, float32 maps, constant border) with a EuRoC-like
                    radial-tangential undistortion map and with a jittered map on a small image, and
                    createCLAHE(3.0,(8,8)) / (2.0,(4,4)) on frame 1 and on sizes that are not multiples
                    of the grid (sha256 for the full frames, arrays for the small ones).
  cv2_hamming.npz   cv2.BFMatcher(NORM_HAMMING).knnMatch (k=3) on seeded random 256-bit codes
                    with planted exact duplicates (tie cases).

    python tests/golden/make_golden.py
"""
import hashlib
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/line_matching/data/mh04/imgs"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def lsd(img, refine):
    d = cv2.createLineSegmentDetector(refine)
    lines, width, prec, nfa = d.detect(img)
    n = 0 if lines is None else len(lines)
    lines = np.zeros((0, 4), np.float32) if n == 0 else lines.reshape(-1, 4)
    width = np.zeros(0) if n == 0 else np.asarray(width, np.float64).ravel()
    prec = np.zeros(0) if n == 0 else np.asarray(prec, np.float64).ravel()
    nfa = np.zeros(0) if (n == 0 or nfa is None) else np.asarray(nfa, np.float64).ravel()
    return lines, width, prec, nfa


def main():
    assert cv2.__version__.startswith("4.13"), cv2.__version__
    frames = np.stack([cv2.imread(os.path.join(REF, f"{k}.png"), cv2.IMREAD_GRAYSCALE) for k in range(1, 16)])
    assert frames.shape == (15, 480, 752) and frames.dtype == np.uint8
    np.savez_compressed(os.path.join(HERE, "mh04_frames.npz"), frames=frames)

    out = {}
    for k in (1, 5, 10):
        g = cv2.GaussianBlur(frames[k - 1], (5, 5), 1)
        lines, width, prec, nfa = lsd(g, cv2.LSD_REFINE_ADV)
        out[f"adv{k}_lines"], out[f"adv{k}_width"], out[f"adv{k}_prec"], out[f"adv{k}_nfa"] = lines, width, prec, nfa
    g = cv2.GaussianBlur(frames[0], (5, 5), 1)
    out["std1_lines"], out["std1_width"], _, _ = lsd(g, cv2.LSD_REFINE_STD)
    out["none1_lines"], out["none1_width"], _, _ = lsd(g, cv2.LSD_REFINE_NONE)
    # octave-1 image of frame 5 (what a 2-octave LSDDetector feeds LSD)
    g5 = cv2.GaussianBlur(frames[4], (5, 5), 1)
    p5 = cv2.pyrDown(g5, dstsize=(752 // 2, 480 // 2))
    out["adv5_oct1_lines"], out["adv5_oct1_width"], _, out["adv5_oct1_nfa"] = lsd(p5, cv2.LSD_REFINE_ADV)
    # known answers: 200x200 step edge 50|200 at column 100, contrast 5 / 6
    step = np.full((200, 200), 50, np.uint8); step[:, 100:] = 200
    out["step_lines"], out["step_width"], _, out["step_nfa"] = lsd(step, cv2.LSD_REFINE_ADV)
    for c in (5, 6):
        s = np.full((200, 200), 100, np.uint8); s[:, 100:] = 100 + c
        out[f"contrast{c}_lines"] = lsd(s, cv2.LSD_REFINE_ADV)[0]
    np.savez_compressed(os.path.join(HERE, "cv2_lsd.npz"), **out)

    rng = np.random.default_rng(1234)
    img = rng.integers(0, 256, (61, 83), dtype=np.uint8)
    pr = {"img": img,
          "blur5": cv2.GaussianBlur(img, (5, 5), 1), "blur7": cv2.GaussianBlur(img, (7, 7), 0.75),
          "resize08": cv2.resize(img, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT),
          "pyrdown": cv2.pyrDown(img, dstsize=(83 // 2, 61 // 2)),
          "sobel_dx": cv2.Sobel(img, cv2.CV_16S, 1, 0, ksize=3), "sobel_dy": cv2.Sobel(img, cv2.CV_16S, 0, 1, ksize=3)}
    f1 = frames[0]
    pr["f1_sha"] = np.array([sha(cv2.GaussianBlur(f1, (5, 5), 1)), sha(cv2.GaussianBlur(f1, (7, 7), 0.75)),
                             sha(cv2.resize(f1, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT)),
                             sha(cv2.pyrDown(f1, dstsize=(376, 240))), sha(cv2.Sobel(f1, cv2.CV_16S, 1, 0, ksize=3)),
                             sha(cv2.Sobel(f1, cv2.CV_16S, 0, 1, ksize=3))])
    yx = rng.integers(-1020, 1021, (4096, 2)).astype(np.float32)
    pr["atan_yx"] = yx
    pr["atan_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(HERE, "cv2_prims.npz"), **pr)

    # ---- pre-processing
    h, w = 480, 752
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375        # EuRoC cam0 intrinsics (config/euroc)
    k1, k2, p1, p2 = -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05
    x = (xx - cx) / fx; y = (yy - cy) / fy; r2 = x * x + y * y; rad = 1 + k1 * r2 + k2 * r2 * r2
    mapx = ((x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)) * fx + cx).astype(np.float32)
    mapy = ((y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y) * fy + cy).astype(np.float32)
    und = cv2.remap(f1, mapx, mapy, cv2.INTER_LINEAR)
    pp = {"euroc_mapx": mapx.astype(np.float16).astype(np.float32), "f1_remap_sha": np.array([sha(und)]),
          "f1_clahe_sha": np.array([sha(cv2.createCLAHE(3.0, (8, 8)).apply(f1))]),
          "f1_remap_clahe_sha": np.array([sha(cv2.createCLAHE(3.0, (8, 8)).apply(und))])}
    del pp["euroc_mapx"]  # the map is regenerated from the formula above in the tests
    rng2 = np.random.default_rng(4321)  # own stream: the vectors above/below do not move
    small = rng2.integers(0, 256, (75, 100), dtype=np.uint8)
    smx = (np.arange(100)[None, :] + rng2.uniform(-4, 4, (75, 100))).astype(np.float32)
    smy = (np.arange(75)[:, None] + rng2.uniform(-4, 4, (75, 100))).astype(np.float32)
    smx[::7, ::5] = np.round(smx[::7, ::5]); smy[::7, ::5] = np.round(smy[::7, ::5])   # exact-integer coordinates
    pp.update(small=small, smx=smx, smy=smy, small_remap=cv2.remap(small, smx, smy, cv2.INTER_LINEAR),
              small_clahe=cv2.createCLAHE(3.0, (8, 8)).apply(small),
              small_clahe_2_4=cv2.createCLAHE(2.0, (4, 4)).apply(small))
    np.savez_compressed(os.path.join(HERE, "cv2_preproc.npz"), **pp)

    q = rng.integers(0, 256, (97, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (131, 32), dtype=np.uint8)
    t[7] = t[3]; t[100] = t[3]; t[50] = q[10]; t[51] = q[10]; t[20] = q[10]; q[11] = t[3]
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=3)
    idx = np.array([[x.trainIdx for x in r] for r in m], np.int32)
    dist = np.array([[x.distance for x in r] for r in m], np.int32)
    np.savez_compressed(os.path.join(HERE, "cv2_hamming.npz"), q=q, t=t, idx=idx, dist=dist)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
