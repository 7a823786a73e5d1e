"""Oracle pinned against cv2 4.13: LSD.  Bit-exact: segment
count, order, float32 endpoints, precision and NFA (i.e. identical pixel counts in every
rectangle).  The float64 width is compared to 1e-13 relative: the oracle takes
cos/sin(theta) from the deterministic correctly rounded sincos (oracle/orc_sincos.h) where
cv2 calls libm, which is not correctly rounded for ~0.1 % of arguments (last-bit only)."""
import numpy as np
import pytest


@pytest.mark.parametrize("k", [1, 5, 10])
def test_lsd_adv_golden_frames(orc, golden, mh04, k):
    g = golden["cv2_lsd"]
    blurred = orc.gaussian_blur5(mh04[k - 1])
    seg, width, prec, nfa = orc.lsd_detect(blurred, refine=2)
    assert seg.shape == g[f"adv{k}_lines"].shape
    assert np.array_equal(seg, g[f"adv{k}_lines"])
    assert np.allclose(width, g[f"adv{k}_width"], rtol=1e-13, atol=0)
    assert np.array_equal(prec, g[f"adv{k}_prec"])
    assert np.array_equal(nfa, g[f"adv{k}_nfa"])


def test_lsd_std_none_golden(orc, golden, mh04):
    g = golden["cv2_lsd"]
    blurred = orc.gaussian_blur5(mh04[0])
    seg, width, _, _ = orc.lsd_detect(blurred, refine=1)
    assert np.array_equal(seg, g["std1_lines"]) and np.allclose(width, g["std1_width"], rtol=1e-13, atol=0)
    seg, width, _, _ = orc.lsd_detect(blurred, refine=0)
    assert np.array_equal(seg, g["none1_lines"]) and np.allclose(width, g["none1_width"], rtol=1e-13, atol=0)


def test_lsd_second_octave_golden(orc, golden, mh04):
    g = golden["cv2_lsd"]
    p = orc.pyrdown(orc.gaussian_blur5(mh04[4]))
    seg, width, _, nfa = orc.lsd_detect(p, refine=2)
    assert np.array_equal(seg, g["adv5_oct1_lines"])
    assert np.allclose(width, g["adv5_oct1_width"], rtol=1e-13, atol=0) and np.array_equal(nfa, g["adv5_oct1_nfa"])


def test_lsd_known_answers(orc, golden):
    g = golden["cv2_lsd"]
    step = np.full((200, 200), 50, np.uint8); step[:, 100:] = 200
    seg, width, _, nfa = orc.lsd_detect(step, refine=2)
    assert np.array_equal(seg, g["step_lines"]) and np.array_equal(nfa, g["step_nfa"])
    assert len(seg) == 1
    # [99.375, 0.625, 99.375, 198.125] at scale 0.8
    assert np.allclose(seg[0], [99.375, 0.625, 99.375, 198.125], atol=1e-3) or \
        np.allclose(seg[0], [99.375, 198.125, 99.375, 0.625], atol=1e-3)
    for c in (5, 6):
        s = np.full((200, 200), 100, np.uint8); s[:, 100:] = 100 + c
        seg = orc.lsd_detect(s, refine=2)[0]
        assert np.array_equal(seg, g[f"contrast{c}_lines"])
    assert len(g["contrast5_lines"]) == 0  # below the gradient threshold rho = 5.226 after the 0.8 pre-scaling


def test_lsd_degenerate_images(orc):
    assert len(orc.lsd_detect(np.zeros((64, 64), np.uint8))[0]) == 0
    assert len(orc.lsd_detect(np.full((40, 50), 255, np.uint8))[0]) == 0


@pytest.mark.parametrize("seed", [0, 1])
def test_lsd_live_cv2_synthetic(orc, synth, seed):
    cv2 = pytest.importorskip("cv2")
    img = synth.sequence(1, w=376, h=240, seed=100 + seed, n_quads=12, n_strokes=20)[0]
    lines, width, prec, nfa = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV).detect(img)
    seg, w2, p2, n2 = orc.lsd_detect(img, refine=2)
    assert np.array_equal(seg, lines.reshape(-1, 4))
    assert np.allclose(w2, width.ravel(), rtol=1e-13, atol=0)
    # NFA values are equal unless libm's last-bit error on cos/sin(theta) moved one pixel in or
    # out of a rectangle in cv2 (see module docstring): allow max(1, 1 %) segments to differ.
    assert (n2 != nfa.ravel()).sum() <= max(1, len(n2) // 100)
