"""Debug: GPU line matcher vs golden, per-anchor mismatch statistics."""
import importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib
vpl = importlib.import_module("vplines_slam_b200")
spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_golden_linematch.py"))
mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_linematch.npz"))
ctx = vpl.Context(max_width=752, max_height=480, max_lines=512, max_batch=8, lsd_path=False)
for name, (a, b, p, illum, topo) in mk.cases().items():
    ctx.linematch_configure(vpl.capi.LineMatchParam(illumination_adapt=int(illum), topological_filter=int(topo)))
    la = np.ascontiguousarray(gold[name + "_lines_ref"]).view(vpl.capi.LINE_DTYPE).reshape(-1)
    lb = np.ascontiguousarray(gold[name + "_lines_cur"]).view(vpl.capi.LINE_DTYPE).reshape(-1)
    r2c = ctx.linematch_batch([a], [b], [la], [lb])[0]
    d = ctx.linematch_points(0)
    g = {k: gold[name + "_" + k] for k in ("kps_ref", "kps_cur", "status", "err", "kp2line")}
    bad = np.nonzero((d["kps_cur"].view(np.uint32) != g["kps_cur"].view(np.uint32)).any(axis=1))[0]
    print(name, "n", len(g["status"]), "kps_ref ok", np.array_equal(d["kps_ref"], g["kps_ref"]), "bad kps_cur", len(bad),
          "status mism", int((d["status"] != g["status"]).sum()), "err mism", int((d["err"].view(np.uint32) != g["err"].view(np.uint32)).sum()),
          "r2c ok", np.array_equal(r2c, gold[name + "_ref_to_cur"]))
    for i in bad[:6]:
        print("   ", i, d["kps_ref"][i], "gpu", d["kps_cur"][i], "ref", g["kps_cur"][i], "st", d["status"][i], g["status"][i], "diff", d["kps_cur"][i] - g["kps_cur"][i])
