import importlib, os, sys, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
v = importlib.import_module("vplines_slam_b200")
from oracle import oracle as O
mh04 = np.load(os.path.join(ROOT, "tests/golden/mh04_frames.npz"))["frames"]
ctx = v.Context(max_width=752, max_height=480, max_octaves=1, max_lines=4096, max_batch=2, num_slots=1)
L = v.capi.load()
img = O.gaussian_blur5(mh04[3])
for cand in (559, 852):
    L.vpl_debug_set_nfa_cand(cand)
    sys.stdout.flush()
    got = ctx.lsd_raw(img)
    ctx.sync()
    print("=== end cand", cand, flush=True)
