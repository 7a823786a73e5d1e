"""Stage-by-stage GPU-vs-oracle report (more informative than pytest -x on a fresh box)."""
import importlib
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
v = importlib.import_module("vplines_slam_b200")
synth = importlib.import_module("vplines-slam_b200.synth")
from oracle import oracle as O  # noqa: E402

mh04 = np.load(os.path.join(ROOT, "tests/golden/mh04_frames.npz"))["frames"]


def section(name, fn):
    t = time.time()
    try:
        fn()
        print(f"[ OK ] {name}  ({time.time() - t:.2f}s)", flush=True)
    except Exception:
        print(f"[FAIL] {name}\n{traceback.format_exc()}", flush=True)


def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print(f"   {name}: SHAPE {a.shape} vs {b.shape}")
        return False
    ne = a != b
    if ne.any():
        w = np.argwhere(ne)
        print(f"   {name}: {ne.sum()} / {a.size} differ, first at {w[0]}: got {a[tuple(w[0])]} exp {b[tuple(w[0])]}")
        return False
    print(f"   {name}: identical ({a.size})")
    return True


ctx = v.Context(max_width=1280, max_height=720, max_octaves=2, max_lines=4096, max_batch=8, num_slots=2, profile=True)


def prims():
    for shape in [(480, 752), (61, 83), (100, 36)]:
        img = np.random.default_rng(shape[0]).integers(0, 256, shape, dtype=np.uint8)
        cmp(f"blur5 {shape}", ctx.debug_stage(0, img), O.gaussian_blur5(img))
        cmp(f"pyrdown {shape}", ctx.debug_stage(1, img), O.pyrdown(img))
        dx, dy = ctx.debug_stage(2, img); odx, ody = O.sobel3(img)
        cmp(f"sobel dx {shape}", dx, odx); cmp(f"sobel dy {shape}", dy, ody)
        cmp(f"scale08 {shape}", ctx.debug_stage(3, img), O.resize08(O.gaussian_blur7(img)))


def stages():
    img = mh04[0]
    scaled, ang, order = O.lsd_stages(img)
    cmp("scaled", ctx.debug_stage(3, img), scaled)
    cmp("angle", ctx.debug_stage(4, img), ang)
    cmp("order", ctx.debug_stage(5, img), order)


def lsd_raw():
    for k in (1, 5):
        b = O.gaussian_blur5(mh04[k - 1])
        t = time.time(); got = ctx.lsd_raw(b); dt = time.time() - t
        seg, width, prec, nfa = O.lsd_detect(b, refine=2)
        g = np.stack([got["x1"], got["y1"], got["x2"], got["y2"]], 1)
        print(f"   frame {k}: gpu {len(g)} segs, oracle {len(seg)}, {dt * 1e3:.1f} ms")
        if len(g) == len(seg):
            cmp("segments", g, seg)
            print("   max |width diff|", np.abs(got["width"] - width).max(), " max |nfa diff|", np.abs(got["nfa"] - nfa).max())
        else:
            n = min(len(g), len(seg))
            ne = np.argwhere((g[:n] != seg[:n]).any(1))
            print("   first differing row", ne[0] if len(ne) else None)
            if len(ne):
                i = int(ne[0]); print("   got", g[max(0, i - 1):i + 2], "\n   exp", seg[max(0, i - 1):i + 2])


def keylines():
    for octs in (1, 2):
        got = ctx.lsd_detect_batch(mh04[3:6], scale=2, num_octaves=octs)
        for f in range(3):
            exp = O.lsd_detector_detect(mh04[3 + f], 2, octs)
            ok = len(got[f]) == len(exp) and all(np.array_equal(got[f][n], exp[n]) for n in exp.dtype.names)
            print(f"   octaves={octs} frame {f}: {len(got[f])} vs {len(exp)} keylines, identical={ok}")
            if not ok and len(got[f]) == len(exp):
                for n in exp.dtype.names:
                    if not np.array_equal(got[f][n], exp[n]):
                        print("     field", n, "differs in", (got[f][n] != exp[n]).sum())


def lbd():
    frames = mh04[0:2]
    kls = [O.lsd_detector_detect(img, 2, 2) for img in frames]
    got = ctx.lbd_compute_batch(frames, kls)
    for f in range(2):
        exp = O.lbd_compute(frames[f], kls[f])
        cmp(f"lbd frame {f}", got[f], exp)
        bad = np.argwhere((got[f] != exp).any(1)).ravel()
        if len(bad):
            print("     bad lines:", bad[:10], "octaves", kls[f]["octave"][bad[:10]], "npix", kls[f]["numOfPixels"][bad[:10]])


def hamming():
    rng = np.random.default_rng(9)
    for (nq, nt, hi) in [(97, 131, 256), (300, 513, 4), (1, 700, 256), (33, 2, 256), (2000, 2000, 256)]:
        q = rng.integers(0, hi, (nq, 32), dtype=np.uint8); t = rng.integers(0, hi, (nt, 32), dtype=np.uint8)
        for k in (1, 2, 3):
            c2 = ctx if nq <= 4096 else None
            m = ctx.match_batch([q], [t], k=k)[0]
            idx, dist = O.hamming_knn(q, t, k)
            ok = np.array_equal(m["trainIdx"], idx) and np.array_equal(m["distance"][idx >= 0].astype(np.int32), dist[idx >= 0])
            print(f"   {nq}x{nt} k={k}: identical={ok}")


def fused():
    frames = mh04[:10]
    fe = v.FrontEnd(ctx, scale=2, num_octaves=1, k=2)
    kls, descs, ms = fe.run(frames)
    prev = None
    for f, img in enumerate(frames):
        ekl = O.lsd_detector_detect(img, 2, 1); ed = O.lbd_compute(img, ekl)
        okk = len(ekl) == len(kls[f]) and all(np.array_equal(ekl[n], kls[f][n]) for n in ekl.dtype.names)
        okd = descs[f].shape == ed.shape and np.array_equal(descs[f], ed)
        okm = True
        if prev is not None:
            idx, dist = O.hamming_knn(ed, prev, 2)
            okm = ms[f]["trainIdx"].shape == idx.shape and np.array_equal(ms[f]["trainIdx"], idx)
        else:
            okm = bool((ms[f]["trainIdx"] == -1).all())
        print(f"   frame {f}: keylines={okk} desc={okd} match={okm}")
        prev = ed


def timing():
    for name, B, octs in (("C2_euroc_752x480", 64, 1), ("C2_euroc_752x480", 256, 1)):
        c = v.capi.CONFIGS if False else None
        cfg = synth.CONFIGS[name]
        frames = synth.config_sequence(name, 16)
        frames = np.concatenate([frames] * (B // 16))
        c2 = v.Context(max_width=cfg["w"], max_height=cfg["h"], max_octaves=octs, max_lines=1024, max_batch=B,
                       num_slots=2, profile=True)
        for it in range(3):
            c2.reset_stage_times()
            t = time.time()
            kls, descs, ms = c2.frontend_batch(frames, num_octaves=octs, k=1, cap=1024)
            dt = time.time() - t
            st = c2.stage_times()
            print(f"   {name} B={B} it={it}: {dt * 1e3:.1f} ms wall = {B / dt:.0f} fps; lines/frame {np.mean([len(k) for k in kls]):.0f}")
            print("     " + "  ".join(f"{s}={ms_:.2f}" for s, (ms_, n) in st.items()))
        c2.close()


section("image primitives", prims)
section("lsd stages", stages)
section("lsd raw", lsd_raw)
section("keylines", keylines)
section("lbd", lbd)
section("hamming", hamming)
section("fused", fused)
section("timing", timing)
