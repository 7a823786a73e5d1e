import sys, numpy as np, importlib
sys.path.insert(0, '.')
vpl = importlib.import_module("vplines_slam_b200")
rng = np.random.default_rng(1)
img = rng.integers(0, 256, (480, 752), dtype=np.uint8)
with vpl.Context(max_width=752, max_height=480, max_octaves=1, max_lines=256, max_batch=2, num_slots=1) as c:
    out = c.debug_stage(0, img)
    print("ok", out.shape, int(out.sum()))
