"""Generates tests/golden/ref_linematch.npz: outputs of the REFERENCE'S OWN line matcher
(/root/reference/line_matching/src/line_matching.cpp + lk_tracker_invoker_2d.cpp compiled against
oracle/cvshim into oracle/_ref/libref_linefront.so, `make -C oracle ref`) -- run in the authoring
container, where /root/reference exists.  Tests read only the .npz (inputs are regenerated from
the committed frames / seeded generators; the line sets fed to the matcher are stored).

Cases: LineMatching::Matching(prev, cur, lines_prev, lines_cur, ..., illumination_adapt, topological_filter)
  mh04_k      frames k -> k+1 of the bundled EuRoC MH_04 sequence, k = 1, 5, 9, 14, tracker settings
              (true, true), lines = the reference's EDLines with the node parameters
  mh04_far    frames 5 -> 10 (the pair of line_matching/data/line_matching_result.png)
  mh04_plain  frames 3 -> 4 with illumination_adapt = false, topological_filter = false
  synth       two consecutive seeded synthetic frames 376x240 (pyramid stops at level 3: 47x30)
  small       160x100 crop pair (pyramid stops early: 20x13 would be <= the 13-px window)
Stored per case: lines_ref, lines_cur (inputs), ref_to_cur, kps_ref, kps_cur, status, err, kp2line.

    python tests/golden/make_golden_linematch.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def cases():
    synth = importlib.import_module("vplines-slam_b200.synth")
    fr = np.load(os.path.join(HERE, "mh04_frames.npz"))["frames"]
    c = {}
    node = O.EDLineParam()
    for k in (1, 5, 9, 14):
        c[f"mh04_{k}"] = (fr[k - 1], fr[k], node, True, True)
    c["mh04_far"] = (fr[4], fr[9], node, True, True)
    c["mh04_plain"] = (fr[2], fr[3], node, False, False)
    s = synth.sequence(2, w=376, h=240, seed=41, n_quads=12, n_strokes=20)
    c["synth"] = (s[0], s[1], O.EDLineParam(minLineLen=20), True, True)
    c["small"] = (np.ascontiguousarray(fr[6][200:300, 300:460]), np.ascontiguousarray(fr[7][200:300, 300:460]),
                  O.EDLineParam(minLineLen=15), True, True)
    return c


def main():
    assert O.build_ref(), "needs /root/reference"
    out = {}
    for name, (a, b, p, illum, topo) in cases().items():
        la = O.ref_edline_detect(a, p, True)
        lb = O.ref_edline_detect(b, p, True)
        r2c, d = O.ref_line_matching(a, b, la, lb, illum, topo, details=True)
        out[name + "_lines_ref"] = la
        out[name + "_lines_cur"] = lb
        out[name + "_ref_to_cur"] = r2c
        for k, v in d.items():
            out[name + "_" + k] = v
        print(name, a.shape, len(la), len(lb), "matches", int((r2c >= 0).sum()), "anchors", len(d["status"]),
              "tracked", int(d["status"].sum()))
    np.savez_compressed(os.path.join(HERE, "ref_linematch.npz"), **out)


if __name__ == "__main__":
    main()
