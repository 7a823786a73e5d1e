"""Find where GPU LSD and the oracle part ways on a frame (candidate lists)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
v = importlib.import_module("vplines_slam_b200")
from oracle import oracle as O
mh04 = np.load(os.path.join(ROOT, "tests/golden/mh04_frames.npz"))["frames"]
ctx = v.Context(max_width=752, max_height=480, max_octaves=1, max_lines=4096, max_batch=2, num_slots=1)
np.set_printoptions(precision=17, linewidth=200)
imgs = {"f4": O.gaussian_blur5(mh04[3]), "f5o1": O.pyrdown(O.gaussian_blur5(mh04[4])), "f1": O.gaussian_blur5(mh04[0])}
for i in range(15):
    imgs[f"all{i}"] = O.gaussian_blur5(mh04[i])
for name, img in imgs.items():
    got = ctx.lsd_raw(img)
    gc = ctx.debug_candidates()
    oc = O.lsd_candidates(img)
    seg = O.lsd_detect(img, refine=2)
    print(f"== {name}: gpu {len(got)} segs / {len(gc)} cands; oracle {len(seg[0])} segs / {len(oc)} cands")
    n = min(len(gc), len(oc))
    acc_diff = np.argwhere(gc[:n, 13] != oc[:n, 13]).ravel()
    # pre-NFA geometry can only be compared for fields the NFA stage does not touch: x y theta dx dy
    geo = np.abs(gc[:n, 5:10] - oc[:n, 5:10]).max(1)
    bad_geo = np.argwhere(geo > 1e-9).ravel()
    print("   accepted flags differ at", acc_diff[:10], " geometry differs at", bad_geo[:10], " max geo diff", geo.max() if n else None)
    nd = np.abs(gc[:n, 12] - oc[:n, 12])
    big = np.argsort(-nd)[:5]
    for b in big:
        print(f"   cand {b}: nfa gpu {gc[b, 12]:.6f} oracle {oc[b, 12]:.6f} acc {gc[b, 13]}/{oc[b, 13]} size {oc[b, 14]:.0f} width g/o {gc[b, 4]:.4f}/{oc[b, 4]:.4f} p g/o {gc[b, 11]}/{oc[b, 11]}")
    for b in list(acc_diff[:3]) + list(bad_geo[:3]):
        print("   --- cand", b, "\n   gpu", gc[b], "\n   orc", oc[b])
    if name.startswith("all") and len(got) == len(seg[0]) and not len(acc_diff) and not len(bad_geo):
        continue
