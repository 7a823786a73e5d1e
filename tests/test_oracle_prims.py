"""Oracle pinned against cv2 4.13: image primitives."""
import hashlib

import numpy as np
import pytest


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_prims_golden_small(orc, golden):
    g = golden["cv2_prims"]
    img = g["img"]
    assert np.array_equal(orc.gaussian_blur5(img), g["blur5"])
    assert np.array_equal(orc.gaussian_blur7(img), g["blur7"])
    assert np.array_equal(orc.resize08(img), g["resize08"])
    assert np.array_equal(orc.pyrdown(img), g["pyrdown"])
    dx, dy = orc.sobel3(img)
    assert np.array_equal(dx, g["sobel_dx"]) and np.array_equal(dy, g["sobel_dy"])


def test_prims_golden_frame_hashes(orc, golden, mh04):
    f1 = mh04[0]
    dx, dy = orc.sobel3(f1)
    got = [sha(orc.gaussian_blur5(f1)), sha(orc.gaussian_blur7(f1)), sha(orc.resize08(f1)), sha(orc.pyrdown(f1)),
           sha(dx), sha(dy)]
    assert got == list(golden["cv2_prims"]["f1_sha"])


def test_fast_atan2_golden(orc, golden):
    g = golden["cv2_prims"]
    got = np.array([orc.fast_atan2(y, x) for y, x in g["atan_yx"]], np.float32)
    assert np.array_equal(got, g["atan_deg"])


@pytest.mark.parametrize("shape", [(7, 9), (17, 33), (64, 101), (480, 752)])
def test_prims_live_cv2(orc, shape):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    h, w = shape
    assert np.array_equal(orc.gaussian_blur5(img), cv2.GaussianBlur(img, (5, 5), 1))
    assert np.array_equal(orc.gaussian_blur7(img), cv2.GaussianBlur(img, (7, 7), 0.75))
    assert np.array_equal(orc.resize08(img), cv2.resize(img, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT))
    assert np.array_equal(orc.pyrdown(img), cv2.pyrDown(img, dstsize=(w // 2, h // 2)))
    dx, dy = orc.sobel3(img)
    assert np.array_equal(dx, cv2.Sobel(img, cv2.CV_16S, 1, 0, ksize=3))
    assert np.array_equal(dy, cv2.Sobel(img, cv2.CV_16S, 0, 1, ksize=3))
