"""tests/golden/ref_tracker.npz: what the REFERENCE'S OWN LineFeatureTracker::readImage
(/root/reference/feature_tracker/src/line_feature_tracker.cpp compiled into oracle/_ref/libref_tracker.so, see
oracle/ref_tracker_glue.cpp) produces for the 15 bundled EuRoC MH_04 frames, run as one sequence on one object with the
EuRoC configuration (undistortion map of cam0, EQUALIZE, max_h_lines = max_v_lines = 25, min_line_length 35,
line_fit_err 1.8) and time(NULL) == seed0 + frame.  Only runs where /root/reference exists.

    python tests/golden/make_golden_tracker.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from test_oracle_preproc import euroc_maps  # noqa: E402

K = (461.6, 460.3, 363.0, 248.1)  # fx, fy, cx, cy of config/euroc/euroc_config.yaml
CFG = dict(equalize=1, max_h=25, max_v=25, min_len=35.0, fit_err=1.8)
SEEDS = (40, 977)  # sequences on which the reference's vanishing-point stage never reads lx[] out of range


def main():
    frames = np.load(os.path.join(ROOT, "tests", "golden", "mh04_frames.npz"))["frames"]
    mapx, mapy = euroc_maps()
    out = {"K": np.array(K, np.float32), "cfg": np.array([CFG["equalize"], CFG["max_h"], CFG["max_v"]], np.int32),
           "cfg_f": np.array([CFG["min_len"], CFG["fit_err"]], np.float32), "seeds": np.array(SEEDS, np.uint32)}
    for s0 in SEEDS:
        t = O.RefTracker(mapx, mapy, *K, bool(CFG["equalize"]), CFG["max_h"], CFG["max_v"], CFG["min_len"], CFG["fit_err"])
        for i, f in enumerate(frames):
            r = t.read(f, s0 + i)
            p = "s%d_f%02d_" % (s0, i)
            out[p + "lines"] = r["lines"]; out[p + "ids"] = np.array(r["ids"], np.int32)
            out[p + "vps"] = r["vps"]; out[p + "t_cnt"] = np.array(r["t_cnt"], np.int32)
            out[p + "exit"] = np.array([r["lines_exit"]], np.int32)
        t.close()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_tracker.npz"), **out)
    print("wrote ref_tracker.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
