"""Oracle for SURVEY 8f-1 (EDLines, oracle/orc_edlines.c) pinned against the reference's OWN code:
tests/golden/ref_edlines.npz was produced by /root/reference/line_matching/src/edline_detector.cpp
compiled against oracle/cvshim (tests/golden/make_golden_edlines.py); where that build exists
(oracle/_ref/libref_linefront.so) it is also called live.  Bar: bit-exact -- edge-chain pixels,
chain starts and every byte of every Line record, in single-thread order."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_edlines", os.path.join(HERE, "golden", "make_golden_edlines.py"))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)
CASES = mk.cases()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_edlines.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_edlines_golden(orc, gold, name):
    img, p, sm = CASES[name]
    lines, xy, sid = orc.edline_detect(img, p, sm, stages=True)
    assert np.array_equal(xy, gold[name + "_xy"])
    assert np.array_equal(sid, gold[name + "_sid"])
    assert lines.tobytes() == gold[name + "_lines"].tobytes()


def test_edlines_live_reference_build(orc, mh04, synth):
    if not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_linefront.so")):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    imgs = [(mh04[k], orc.EDLineParam(), True) for k in (2, 7, 12)]
    imgs += [(f, orc.EDLineParam(minLineLen=18), False) for f in synth.sequence(3, w=320, h=200, seed=77, n_quads=10, n_strokes=16)]
    n = 0
    for img, p, sm in imgs:
        a = orc.edline_detect(img, p, sm, stages=True)
        b = orc.ref_edline_detect(img, p, sm, stages=True)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert a[0].tobytes() == b[0].tobytes()
        n += len(a[0])
    assert n > 100


def test_edlines_known_answer(orc):
    # a sharp 50|200 step has two equal gradient columns -> no anchor (>= neighbour + 5 fails on the
    # plateau) -> EdgeDrawing fails -> no lines.  With a one-pixel ramp (50, 125, 200) the gradient
    # peaks on the ramp column 101 (odd columns are the scanned ones): lines lie exactly on x = 101.
    img = np.full((200, 200), 50, np.uint8); img[:, 100:] = 200
    assert len(orc.edline_detect(img, orc.EDLineParam(minLineLen=25), True)) == 0
    img = np.full((200, 200), 50, np.uint8); img[:, 101] = 125; img[:, 102:] = 200
    lines = orc.edline_detect(img, orc.EDLineParam(minLineLen=25), True)
    assert len(lines) >= 1
    for l in lines:
        assert l["endpoint"][0] == 101.0 and l["endpoint"][2] == 101.0
        assert abs(l["equation"][0]) == 1.0 and l["equation"][1] == 0.0 and abs(l["equation"][2]) == 101.0
    assert sum(l["length"] for l in lines) > 150
    # no gradient above the threshold -> EdgeDrawing fails ("lines not found") -> no lines
    assert len(orc.edline_detect(np.full((64, 80), 9, np.uint8))) == 0


def test_edlines_capacity_error_is_no_lines(orc):
    # dense noise scanned at every pixel: more anchors than W*H/5 -> the reference returns -1
    # (edline_detector.cpp:166-169) after writing out of bounds; the oracle reports no lines.
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (96, 128), dtype=np.uint8)
    lines, xy, sid = orc.edline_detect(img, orc.EDLineParam(scanIntervals=1, anchorThreshold=0, gradientThreshold=0, minLineLen=5), True, stages=True)
    assert len(lines) == 0 and len(xy) == 0


def test_edlines_nfa_matches_lsd_formula(orc):
    import ctypes
    L = orc.lib(); L.orc_ed_nfa.restype = ctypes.c_double
    L.orc_ed_nfa.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double]
    logNT = 2.0 * (np.log10(752.0) + np.log10(480.0))
    assert L.orc_ed_nfa(0, 0, 0.125, logNT) == -logNT
    assert L.orc_ed_nfa(40, 40, 0.125, logNT) == -logNT - 40 * np.log10(0.125)
    v = [L.orc_ed_nfa(100, k, 0.125, logNT) for k in (10, 30, 60, 90)]
    assert all(b > a for a, b in zip(v, v[1:])) and v[0] < 0 < v[2]
