"""the oracle's restatement of LineFeatureTracker::readImage's bookkeeping (oracle/orc_tracker.py) against
(1) tests/golden/ref_tracker.npz, produced by the reference's own line_feature_tracker.cpp, and (2) that code itself
(oracle/_ref/libref_tracker.so) where it is available.  Bar: every Line byte, every id, every track count (except the
one entry the reference reads past the end of its array), vanishing-point vectors bit-exact with libm (math_mode 0)
and within 1e-15 with the shared correctly rounded functions (math_mode 1, what the device computes)."""
import os

import numpy as np
import pytest

from test_oracle_preproc import euroc_maps

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_tracker.npz"))


def run_oracle(orc, mh04, gold, s0, mode):
    from oracle import orc_tracker as T
    mapx, mapy = euroc_maps()
    fx, fy, cx, cy = (float(v) for v in gold["K"])
    eq, mh, mv = (int(v) for v in gold["cfg"])
    ml, fe = (float(v) for v in gold["cfg_f"])
    t = T.Tracker(mapx, mapy, fx, cx, cy, bool(eq), mh, mv, ml, fe, math_mode=mode)
    return t, [dict(t.read(f, s0 + i)) for i, f in enumerate(mh04)]


def check_frame(got, gold, p, mode, oob):
    assert np.asarray(got["lines"]).tobytes() == gold[p + "lines"].tobytes(), p
    assert list(got["ids"]) == list(gold[p + "ids"]), p
    v = np.asarray(got["vps"], np.float64).reshape(-1, 4)
    g = gold[p + "vps"]
    assert v.shape == g.shape, p
    if mode == 0:
        assert v.tobytes() == g.tobytes(), p
    else:
        assert np.allclose(v, g, rtol=0, atol=1e-15, equal_nan=True), p
    t, gt = list(got["t_cnt"]), list(gold[p + "t_cnt"])
    assert len(t) == len(gt), p
    bad = [i for i in range(len(t)) if t[i] != gt[i]]
    assert set(bad) <= oob, (p, bad)  # only where the reference read past the end of the previous frame's counters


@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_tracker_equals_reference_golden(orc, mh04, gold, mode):
    for s0 in (int(s) for s in gold["seeds"]):
        t, frames = run_oracle(orc, mh04, gold, s0, mode)
        assert not any(n[0] == "vp_lx_out_of_range" for n in t.notes)
        for i, got in enumerate(frames):
            oob = {n[1] for n in t.notes if n[0] == "t_cnt_out_of_range"}
            check_frame(got, gold, "s%d_f%02d_" % (s0, i), mode, oob)
        assert len(frames[0]["vps"]) == 0 and len(frames[1]["vps"]) == len(frames[1]["lines"])
        assert max(len(f["lines"]) for f in frames[1:]) <= 50 + 25  # tracked + the two quotas
        assert any(a in set(frames[i]["ids"]) for i in range(2, 15) for a in frames[i - 1]["ids"])  # ids do propagate


def test_oracle_tracker_equals_reference_live(orc, mh04, gold):
    if not orc.ref_tracker_available():
        pytest.skip("oracle/_ref/libref_tracker.so not present")
    mapx, mapy = euroc_maps()
    fx, fy, cx, cy = (float(v) for v in gold["K"])
    ref = orc.RefTracker(mapx, mapy, fx, fy, cx, cy, True, 25, 25, 35.0, 1.8)
    try:
        t, frames = run_oracle(orc, mh04[:8], gold, 5000, 0)
        for i, got in enumerate(frames):
            r = ref.read(mh04[i], 5000 + i)
            assert r["lines"].tobytes() == np.asarray(got["lines"]).tobytes() and r["ids"] == list(got["ids"]), i
            assert r["lines_exit"] is True
            if not any(n[0] == "vp_lx_out_of_range" for n in t.notes):
                assert r["vps"].tobytes() == np.asarray(got["vps"], np.float64).reshape(-1, 4).tobytes(), i
    finally:
        ref.close()


def test_oracle_tracker_no_lines_frame(orc, gold):
    """A frame without lines: lines_exit = false, the previous frame stays current (line_feature_tracker.cpp:89-94)."""
    from oracle import orc_tracker as T
    mapx, mapy = euroc_maps()
    t = T.Tracker(mapx, mapy, 461.6, 363.0, 248.1)
    flat = np.full((480, 752), 90, np.uint8)
    cur = t.read(flat, 1)
    assert t.lines_exit is False and len(cur["lines"]) == 0
