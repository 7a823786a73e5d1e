"""Oracle pinned against cv2 4.13: the pre-processing in front of the line path -- cv::remap
(INTER_LINEAR, CV_32FC1 maps, constant border) and cv::CLAHE (readImage,
feature_tracker/src/line_feature_tracker.cpp:62-68).  Bit-exact."""
import hashlib

import numpy as np
import pytest


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def euroc_maps(w=752, h=480):
    """Radial-tangential undistortion map of EuRoC cam0 (config/euroc/euroc_config.yaml intrinsics)."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375
    k1, k2, p1, p2 = -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05
    x = (xx - cx) / fx; y = (yy - cy) / fy; r2 = x * x + y * y; rad = 1 + k1 * r2 + k2 * r2 * r2
    mapx = ((x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)) * fx + cx).astype(np.float32)
    mapy = ((y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y) * fy + cy).astype(np.float32)
    return mapx, mapy


def test_preproc_golden_small(orc):
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "cv2_preproc.npz"))
    assert np.array_equal(orc.remap_linear(g["small"], g["smx"], g["smy"]), g["small_remap"])
    assert np.array_equal(orc.clahe(g["small"], 3.0, 8), g["small_clahe"])       # 75x100: not a multiple of 8
    assert np.array_equal(orc.clahe(g["small"], 2.0, 4), g["small_clahe_2_4"])


def test_preproc_golden_frame(orc, mh04):
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "cv2_preproc.npz"))
    mapx, mapy = euroc_maps()
    und = orc.remap_linear(mh04[0], mapx, mapy)
    assert sha(und) == g["f1_remap_sha"][0]
    assert sha(orc.clahe(mh04[0], 3.0, 8)) == g["f1_clahe_sha"][0]
    assert sha(orc.clahe(und, 3.0, 8)) == g["f1_remap_clahe_sha"][0]


@pytest.mark.parametrize("shape", [(60, 80), (64, 101), (37, 64), (480, 752)])
def test_preproc_live_cv2(orc, shape):
    cv2 = pytest.importorskip("cv2")
    h, w = shape
    rng = np.random.default_rng(h * w)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    mx = (np.arange(w)[None, :] + rng.uniform(-6, 6, shape)).astype(np.float32)
    my = (np.arange(h)[:, None] + rng.uniform(-6, 6, shape)).astype(np.float32)
    mx[::3, ::4] = np.round(mx[::3, ::4]); my[::3, ::4] = np.round(my[::3, ::4])
    assert np.array_equal(orc.remap_linear(img, mx, my), cv2.remap(img, mx, my, cv2.INTER_LINEAR))
    for clip, tiles in ((3.0, 8), (2.0, 4), (40.0, 8)):
        assert np.array_equal(orc.clahe(img, clip, tiles), cv2.createCLAHE(clip, (tiles, tiles)).apply(img))
