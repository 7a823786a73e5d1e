"""Oracle: Hamming kNN pinned against cv2.BFMatcher; LSDDetector packing and LBD have no
upstream vectors (PARITY UNPINNED) -- their tests pin the published invariants."""
import numpy as np
import pytest


def test_hamming_golden_bfmatcher(orc, golden):
    g = golden["cv2_hamming"]
    idx, dist = orc.hamming_knn(g["q"], g["t"], 3)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    # planted ties: three exact duplicates come back in index order
    assert list(idx[11]) == [3, 7, 100] and list(dist[11]) == [0, 0, 0]
    assert list(idx[10]) == [20, 50, 51]
    idx1, dist1 = orc.hamming_knn(g["q"], g["t"], 1)
    assert np.array_equal(idx1[:, 0], g["idx"][:, 0])


def test_hamming_live_bfmatcher(orc):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    q = rng.integers(0, 4, (200, 32), dtype=np.uint8)  # few bits set -> many distance ties
    t = rng.integers(0, 4, (300, 32), dtype=np.uint8)
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    idx, dist = orc.hamming_knn(q, t, 2)
    assert np.array_equal(idx, np.array([[x.trainIdx for x in r] for r in m]))
    assert np.array_equal(dist, np.array([[int(x.distance) for x in r] for r in m]))


def test_hamming_edge_cases(orc):
    q = np.zeros((3, 32), np.uint8)
    idx, dist = orc.hamming_knn(q, np.zeros((0, 32), np.uint8), 2)
    assert (idx == -1).all()
    idx, dist = orc.hamming_knn(q, np.full((1, 32), 255, np.uint8), 2)
    assert list(idx[0]) == [0, -1] and dist[0, 0] == 256


def test_keyline_packing_invariants(orc, mh04):
    img = mh04[4]
    kl = orc.lsd_detector_detect(img, scale=2, num_octaves=2)
    assert len(kl) > 500 and set(np.unique(kl["octave"])) == {0, 1}
    assert np.array_equal(kl["class_id"], np.arange(len(kl)))
    # octave-0 keylines are exactly the LSD segments of the blurred image, clamped to the frame
    seg = orc.lsd_detect(orc.gaussian_blur5(img), refine=2)[0]
    k0 = kl[kl["octave"] == 0]
    assert len(k0) == len(seg)
    exp = seg.copy()
    # checkLineExtremes: < 0 -> 0, >= cols -> cols - 1 (values in (cols-1, cols) are left alone)
    for cols, ix in ((752, [0, 2]), (480, [1, 3])):
        v = exp[:, ix]
        v[v < 0] = 0
        v[v >= cols] = cols - 1
        exp[:, ix] = v
    got = np.stack([k0["startPointX"], k0["startPointY"], k0["endPointX"], k0["endPointY"]], 1)
    assert np.array_equal(got, exp)
    k1 = kl[kl["octave"] == 1]
    assert np.array_equal(k1["startPointX"], k1["sPointInOctaveX"] * 2)
    assert np.array_equal(k1["endPointY"], k1["ePointInOctaveY"] * 2)
    dx = np.abs(np.rint(kl["ePointInOctaveX"]) - np.rint(kl["sPointInOctaveX"]))
    dy = np.abs(np.rint(kl["ePointInOctaveY"]) - np.rint(kl["sPointInOctaveY"]))
    assert np.array_equal(kl["numOfPixels"], (np.maximum(dx, dy) + 1).astype(np.int32))
    ln = np.hypot(kl["sPointInOctaveX"] - kl["ePointInOctaveX"], kl["sPointInOctaveY"] - kl["ePointInOctaveY"])
    assert np.allclose(kl["lineLength"], ln, rtol=1e-6)
    assert np.allclose(kl["angle"], np.arctan2(kl["endPointY"] - kl["startPointY"], kl["endPointX"] - kl["startPointX"]),
                       atol=1e-6)
    # blur_first=0 variant = LSD on the raw image
    kl_raw = orc.lsd_detector_detect(img, scale=2, num_octaves=1, blur_first=False)
    assert len(kl_raw) == len(orc.lsd_detect(img, refine=2)[0])


def test_lbd_invariants(orc, mh04):
    img = mh04[0]
    kl = orc.lsd_detector_detect(img, scale=2, num_octaves=2)
    desc, fd = orc.lbd_compute(img, kl, return_float=True)
    assert desc.shape == (len(kl), 32) and fd.shape == (len(kl), 72)
    assert np.isfinite(fd).all()
    # unit norm after the final renormalisation, entries bounded by the 0.4 clamp / renorm
    assert np.allclose(np.linalg.norm(fd, axis=1), 1.0, atol=1e-5)
    assert (fd >= 0).all()
    # descriptor row order = keyline order (permuting keylines permutes rows)
    perm = np.random.default_rng(0).permutation(len(kl))
    assert np.array_equal(orc.lbd_compute(img, kl[perm]), desc[perm])
    # binarisation: byte c bit i = [desc[8a+i] > desc[8b+i]] for pair c=(0,1)
    bits = (fd[:, 0:8] > fd[:, 8:16]).astype(np.uint8)
    assert np.array_equal(desc[:, 0], (bits * (1 << np.arange(8))).sum(1).astype(np.uint8))
    # descriptors discriminate: consecutive real frames match mostly to nearby lines
    kl2 = orc.lsd_detector_detect(mh04[1], scale=2, num_octaves=1)
    d2 = orc.lbd_compute(mh04[1], kl2)
    k1 = kl[kl["octave"] == 0]; d1 = desc[kl["octave"] == 0]
    idx, dist = orc.hamming_knn(d2, d1, 2)
    good = dist[:, 0] < 0.7 * dist[:, 1]
    shift = np.hypot(kl2["pt_x"][good] - k1["pt_x"][idx[good, 0]], kl2["pt_y"][good] - k1["pt_y"][idx[good, 0]])
    assert good.sum() > 100 and np.median(shift) < 20
