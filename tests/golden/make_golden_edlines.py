"""Generates tests/golden/ref_edlines.npz: outputs of the REFERENCE'S OWN EDLines detector
(/root/reference/line_matching/src/edline_detector.cpp compiled against oracle/cvshim into
oracle/_ref/libref_linefront.so, `make -C oracle ref`) -- run in the authoring container, where
/root/reference exists.  The GPU box has neither; tests read only the .npz.

Cases (Line records in single-thread order + edge-chain pixels + chain starts):
  demo_k     line_matching/src/test_edline_detector.cpp:15 parameters {5,1.0,30,5,2,25,1.8},
             smoothed=false, on mh04 frames k = 1, 5, 10
  node_k     the tracker node's parameters (line_feature_tracker_node.cpp:203 + euroc yaml:
             minLineLen 35, fitErr 1.8), smoothed=true, frames k = 1, 5, 10, 15
  synth_*    seeded synthetic frames (vplines-slam_b200/synth.py): 376x240 and an odd size 331x207
  noise      uniform noise 160x120 (many short chains; kept below the reference's array capacities --
             beyond them the reference writes out of bounds), flat image (EdgeDrawing returns -1)

    python tests/golden/make_golden_edlines.py
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
    for k in (1, 5, 10):
        c[f"demo_{k}"] = (fr[k - 1], O.EDLineParam(minLineLen=25), False)
    for k in (1, 5, 10, 15):
        c[f"node_{k}"] = (fr[k - 1], O.EDLineParam(), True)
    c["synth_a"] = (synth.sequence(1, w=376, h=240, seed=21, n_quads=12, n_strokes=20)[0], O.EDLineParam(minLineLen=20), False)
    odd = synth.sequence(1, w=376, h=240, seed=22, n_quads=12, n_strokes=20)[0][11:218, 17:348]
    c["synth_odd"] = (np.ascontiguousarray(odd), O.EDLineParam(minLineLen=15, lineFitErrThreshold=1.4, anchorThreshold=2, gradientThreshold=20), True)
    rng = np.random.default_rng(5)
    c["noise"] = (rng.integers(0, 256, (120, 160), dtype=np.uint8), O.EDLineParam(minLineLen=12, anchorThreshold=8), False)
    c["flat"] = (np.full((64, 80), 77, np.uint8), O.EDLineParam(), True)
    return c


def main():
    assert O.build_ref(), "needs /root/reference"
    out = {}
    for name, (img, p, sm) in cases().items():
        lines, xy, sid = O.ref_edline_detect(img, p, sm, stages=True)
        out[name + "_lines"] = lines
        out[name + "_xy"] = xy
        out[name + "_sid"] = sid
        print(name, img.shape, len(lines), "lines", len(xy), "edge px", len(sid) - 1 if len(sid) else 0, "chains")
    np.savez_compressed(os.path.join(HERE, "ref_edlines.npz"), **out)


if __name__ == "__main__":
    main()
