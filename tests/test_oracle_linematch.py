"""Oracle for SURVEY 8f-2 (LineMatching::Matching, oracle/orc_linematch.c) pinned against the
reference's OWN code: tests/golden/ref_linematch.npz was produced by the reference's
line_matching.cpp + lk_tracker_invoker_2d.cpp compiled against oracle/cvshim
(tests/golden/make_golden_linematch.py); where that build exists it is also called live.  The
OpenCV calls KLT::calc2D makes itself (pyramid, Scharr, meanStdDev) are pinned against cv2 4.13.
Bar: bit-exact -- tracked positions, status, error, closest-line labels and the match vector."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_linematch", os.path.join(HERE, "golden", "make_golden_linematch.py"))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)
CASES = mk.cases()
HAVE_REF = os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_linefront.so"))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_linematch.npz"))


def same_bits(a, b):
    return a.shape == b.shape and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("name", sorted(CASES))
def test_linematch_golden(orc, gold, name):
    a, b, p, illum, topo = CASES[name]
    # the line sets are the oracle's own EDLines output: must equal what the reference detected
    la, lb = orc.edline_detect(a, p, True), orc.edline_detect(b, p, True)
    assert same_bits(la, gold[name + "_lines_ref"]) and same_bits(lb, gold[name + "_lines_cur"])
    r2c, d = orc.line_matching(a, b, la, lb, illum=illum, topo=topo, details=True)
    assert np.array_equal(r2c, gold[name + "_ref_to_cur"])
    for k in ("kps_ref", "kps_cur", "status", "err", "kp2line"):
        assert same_bits(d[k], gold[name + "_" + k]), k
    assert (r2c >= 0).sum() > 10


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
def test_linematch_live_reference_build(orc, mh04, synth):
    p = orc.EDLineParam()
    pairs = [(mh04[i], mh04[i + 1], p) for i in (2, 7, 11)]
    s = synth.sequence(3, w=320, h=200, seed=78, n_quads=10, n_strokes=16)
    pairs += [(s[0], s[1], orc.EDLineParam(minLineLen=18)), (s[1], s[2], orc.EDLineParam(minLineLen=18))]
    for a, b, pp in pairs:
        la, lb = orc.edline_detect(a, pp, True), orc.edline_detect(b, pp, True)
        r1, d1 = orc.line_matching(a, b, la, lb, details=True)
        r2, d2 = orc.ref_line_matching(a, b, la, lb, details=True)
        assert np.array_equal(r1, r2)
        for k in d1:
            assert same_bits(d1[k], d2[k]), k


def test_linematch_empty_inputs(orc, mh04):
    la = orc.edline_detect(mh04[0])
    none = la[:0]
    assert orc.line_matching(mh04[0], mh04[1], none, la) is None   # Matching returns false, lm.cpp:621
    assert orc.line_matching(mh04[0], mh04[1], la, none) is None


def test_identity_pair_matches_itself(orc, mh04):
    la = orc.edline_detect(mh04[3])
    r2c, d = orc.line_matching(mh04[3], mh04[3], la, la, details=True)
    ok = r2c >= 0
    assert ok.mean() > 0.8 and np.array_equal(r2c[ok], np.nonzero(ok)[0])
    tr = d["status"] == 1
    assert np.abs(d["kps_cur"][tr] - d["kps_ref"][tr]).max() < 1e-3


def test_pyramid_and_scharr_vs_cv2(orc):
    cv2 = pytest.importorskip("cv2")
    import ctypes

    class Lvl(ctypes.Structure):
        _fields_ = [("w", ctypes.c_int), ("h", ctypes.c_int), ("pad", ctypes.c_int), ("stride", ctypes.c_int),
                    ("img", ctypes.POINTER(ctypes.c_uint8)), ("deriv", ctypes.POINTER(ctypes.c_int16))]
    L = orc.lib()
    L.orc_klt_build_levels.restype = ctypes.c_int
    rng = np.random.default_rng(3)
    for shape in ((61, 83), (480, 752), (100, 160), (31, 45)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        top_cv, pyr = cv2.buildOpticalFlowPyramid(img, (13, 13), 3, withDerivatives=False)
        lv = (Lvl * 8)()
        top = L.orc_klt_build_levels(img.ctypes.data_as(ctypes.c_void_p), shape[1], shape[0], 13, 3, 1, lv)
        assert top == top_cv
        for i in range(top + 1):
            w, h, pad, st = lv[i].w, lv[i].h, lv[i].pad, lv[i].stride
            assert (h, w) == pyr[i].shape and pad == 13 and st == w + 26
            buf = np.ctypeslib.as_array(lv[i].img, shape=(h + 26, st))
            exp = cv2.copyMakeBorder(pyr[i], 13, 13, 13, 13, cv2.BORDER_REFLECT_101)
            assert np.array_equal(buf, exp)
            d = np.ctypeslib.as_array(lv[i].deriv, shape=(h + 26, st, 2))
            lvl = np.ascontiguousarray(pyr[i])
            ex = np.zeros((h + 26, st, 2), np.int16)
            ex[13:13 + h, 13:13 + w, 0] = cv2.Scharr(lvl, cv2.CV_16S, 1, 0)
            ex[13:13 + h, 13:13 + w, 1] = cv2.Scharr(lvl, cv2.CV_16S, 0, 1)
            assert np.array_equal(d, ex)
        L.orc_klt_free_levels(lv, top)
    for shape in ((61, 83), (60, 82), (7, 9)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(orc.pyrdown_std(img), cv2.pyrDown(img))
