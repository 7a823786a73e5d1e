"""Generates tests/golden/ref_vp.npz: outputs of the REFERENCE'S OWN vanishing-point stage
(/root/reference/feature_tracker/src/vanishing_point_detection.cpp compiled against oracle/cvshim
into oracle/_ref/libref_vp.so, `make -C oracle ref`, its time(NULL) answered with the stored seed)
-- run in the authoring container, where /root/reference exists.  Tests read only the .npz.

Cases: run_vanishing_point_detection(img, lines, all_lines, vps, local_vp_ids) after init(f, cx, cy, .)
  mh04_k_s      EDLines (tracker-node parameters) of frame k of the bundled EuRoC MH_04 sequence,
                lines == all_lines, EuRoC intrinsics; first call of the object (frame_count 0)
  mh04_k_s_n    the same on an object that has made a call before (frame_count > 0: vps[1]/vps[2] rule)
  vertical_k    `lines` = the near-vertical subset (what readImage passes as verticalLine,
                line_feature_tracker.cpp:241), all_lines = everything
  manhattan     400 synthetic segments drawn towards three vanishing points, 1280x720 intrinsics
  few           4 lines
Only seeds on which the reference does not read lx[] out of range are kept (see oracle/orc_vp.c);
the generator tries the seeds in order and stores the one it used.
Stored per case: lines, all_lines, f, cx, cy, seed, frame_count (inputs); vps (3x3 float64), vp_idx.

    python tests/golden/make_golden_vp.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

EUROC = (461.6, 363.0, 248.1)  # fx, cx, cy of config/euroc/euroc_config.yaml (projection_parameters)


def manhattan_lines(n=400, w=1280, h=720, f=640.0, seed=5):
    """n segments whose supporting lines pass through one of three orthogonal vanishing points (+ noise)."""
    rng = np.random.default_rng(seed)
    a, b = np.deg2rad(12.0), np.deg2rad(-7.0)
    Ra = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Rb = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    R = Ra @ Rb
    vps = [(f * R[0, k] / R[2, k] + w / 2, f * R[1, k] / R[2, k] + h / 2) for k in range(3)]
    out = np.zeros(n, O.LINE_DTYPE)
    for i in range(n):
        vx, vy = vps[i % 3]
        mx, my = rng.uniform(40, w - 40), rng.uniform(40, h - 40)
        d = np.array([vx - mx, vy - my]); d /= np.linalg.norm(d)
        ang = rng.normal(0, 0.004)
        d = np.array([d[0] * np.cos(ang) - d[1] * np.sin(ang), d[0] * np.sin(ang) + d[1] * np.cos(ang)])
        L = rng.uniform(25, 300) / 2
        p, q = np.array([mx, my]) - L * d, np.array([mx, my]) + L * d
        out["endpoint"][i] = (p[0], p[1], q[0], q[1])
        out["center"][i] = (mx, my)
        out["length"][i] = 2 * L
        nrm = np.array([-d[1], d[0]])
        out["equation"][i] = (nrm[0], nrm[1], -(nrm[0] * mx + nrm[1] * my))
    return out


def inputs():
    """name -> (lines, all_lines, f, cx, cy, frame_count, candidate seeds)."""
    fr = np.load(os.path.join(HERE, "mh04_frames.npz"))["frames"]
    c = {}
    seeds = [1700000000, 1700000001, 1700000002] + list(range(1, 400))
    for k in (1, 4, 8, 12, 15):
        ln = O.edline_detect(fr[k - 1])
        c[f"mh04_{k}"] = (ln, ln, *EUROC, 0, [s + 31 * k for s in seeds])
        c[f"mh04_{k}_next"] = (ln, ln, *EUROC, 3, [s + 31 * k + 7 for s in seeds])
    for k in (2, 9):
        ln = O.edline_detect(fr[k - 1])
        e = ln["endpoint"]
        vert = np.abs(e[:, 0] - e[:, 2]) < 0.35 * np.abs(e[:, 1] - e[:, 3])
        assert vert.sum() > 2
        c[f"vertical_{k}"] = (ln[vert], ln, *EUROC, 1, [s + k for s in seeds])
    m = manhattan_lines()
    c["manhattan"] = (m, m, 640.0, 640.0, 360.0, 0, seeds)
    c["manhattan_next"] = (m, m, 640.0, 640.0, 360.0, 1, [s + 100 for s in seeds])
    few = O.edline_detect(fr[0])[[3, 17, 40, 77]]
    c["few"] = (few, few, *EUROC, 0, seeds)
    return c


def main():
    assert O.build_ref(), "needs /root/reference"
    out = {}
    for name, (ln, al, f, cx, cy, fc, seeds) in inputs().items():
        for seed in seeds:
            _, _, d = O.vp_detect(ln, al, f, cx, cy, seed, fc, math_mode=0, details=True)
            if d["flags"] == 0:
                break
        else:
            raise SystemExit("no usable seed for " + name)
        vps, idx = O.ref_vp_detect(ln, al, f, cx, cy, seed, fc)
        out[name + "_lines"] = ln; out[name + "_all_lines"] = al
        out[name + "_cam"] = np.array([f, cx, cy], np.float32)
        out[name + "_seed"] = np.array([seed, fc], np.int64)
        out[name + "_vps"] = vps; out[name + "_vp_idx"] = idx
        print(name, len(ln), len(al), "seed", seed, "labels", np.bincount(idx, minlength=4))
    np.savez_compressed(os.path.join(HERE, "ref_vp.npz"), **out)


if __name__ == "__main__":
    main()
