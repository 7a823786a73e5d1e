"""The header-only C++ facade (compat/line_descriptor.hpp, compat/vplines_seam.hpp): compiles
against the C ABI without OpenCV; on the GPU it must produce what the ctypes path produces."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vplines-slam_b200")


def build_facade(tmp_path, name="test_facade"):
    exe = str(tmp_path / name)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
           "-L", PKG, "-lvplines_b200", f"-Wl,-rpath,{PKG}"]
    subprocess.check_call(cmd)
    return exe


def make_image(w, h, shift):
    y, x = np.mgrid[0:h, 0:w]
    v = 110 + ((x // 7 + y // 11) % 3) * 4
    v = np.where((x + shift > 60) & (x + shift < 200) & (y > 40) & (y < 150), 40, v)
    v = np.where((y > 170) & (y < 178) & (x > 30), 220, v)
    v = np.where(((x + shift) - y > 150) & ((x + shift) - y < 160), 230, v)
    return v.astype(np.uint8)


def test_facade_compiles_and_fails_loudly_without_gpu(vpl, tmp_path):
    exe = build_facade(tmp_path)
    assert subprocess.run([exe, "--compile-only"]).returncode == 0
    if vpl.capi.device_count() == 0:
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU path" in r.stdout


@pytest.mark.gpu
def test_facade_matches_ctypes_path(vpl, orc, tmp_path):
    exe = build_facade(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    a, b = make_image(320, 240, 0), make_image(320, 240, 3)
    ekl = orc.lsd_detector_detect(a, 2, 2)
    edesc = orc.lbd_compute(a, ekl)
    digest = 1469598103934665603
    for v in edesc.reshape(-1):
        digest = ((digest ^ int(v)) * 1099511628211) % (1 << 64)
    m = re.search(r"keylines=(\d+) self_matches=(\d+) desc_digest=(\d+)", out)
    assert m, out
    assert int(m.group(1)) == len(ekl) and int(m.group(3)) == digest
    # every descriptor's nearest neighbour in its own set is at distance 0
    assert int(m.group(2)) == len(ekl)
    la = orc.lsd_detector_detect(a, 2, 1); lb = orc.lsd_detector_detect(b, 2, 1)
    m2 = re.search(r"lines_a=(\d+) lines_b=(\d+) matched=(\d+)", out)
    assert int(m2.group(1)) == len(la) and int(m2.group(2)) == len(lb)
    idx, dist = orc.hamming_knn(orc.lbd_compute(b, lb), orc.lbd_compute(a, la), 1)
    exp = len(set(int(t) for t, d in zip(idx[:, 0], dist[:, 0]) if d < 30))
    assert int(m2.group(3)) == exp
    assert "depth_check=Error, depth image!= 0" in out
    # C++ batch driver: the sharded run (2 ranks, one-frame halo) equals the single run, and both equal the oracle
    mb = re.search(r"batch lines=(\d+) matched=(\d+) digest=(\d+) sharded_equal=(\d)", out)
    assert mb and mb.group(4) == "1", out
    seq = [make_image(320, 240, 2 * i) for i in range(5)]
    kls = [orc.lsd_detector_detect(f, 2, 1) for f in seq]
    ds = [orc.lbd_compute(f, k) for f, k in zip(seq, kls)]
    assert int(mb.group(1)) == sum(len(k) for k in kls)
    exp_matched = sum(len(ds[i]) for i in range(1, 5) if len(ds[i - 1]))
    assert int(mb.group(2)) == exp_matched
    # the grouped fast path (vpl_frontend_upload + vpl_frontend_submit_group over contiguous frames) == the single run
    gg = re.search(r"grouped equal=(\d) sharded_equal=(\d)", out)
    assert gg and gg.group(1) == "1" and gg.group(2) == "1", out
    # in-process multi-GPU driver (std::thread + context per device, ordered host gather) == the single run
    mg = re.search(r"multigpu world=(\d+) equal=(\d) ordered=(\d)", out)
    assert mg and int(mg.group(1)) >= 2 and mg.group(2) == "1" and mg.group(3) == "1", out


def fnv(data):
    d = 1469598103934665603
    for v in data:
        d = ((d ^ int(v)) * 1099511628211) % (1 << 64)
    return d


def line_digest(lines):
    raw = np.ascontiguousarray(lines).view(np.uint8).reshape(len(lines), 56)[:, :52]
    return fnv(raw.reshape(-1))


def test_reference_named_facade_compiles_and_fails_loudly_without_gpu(vpl, tmp_path):
    exe = build_facade(tmp_path, "test_refseam")
    assert subprocess.run([exe, "--compile-only"]).returncode == 0
    if vpl.capi.device_count() == 0:
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU path" in r.stdout


@pytest.mark.gpu
def test_reference_named_facade_matches_oracle(vpl, orc, tmp_path):
    """EDLineDetector::EDline / LineMatching::Matching of compat/line_matching_b200.hpp, used as the
    tracker uses the reference's classes, against the CPU oracle (= the reference's own code)."""
    exe = build_facade(tmp_path, "test_refseam")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    a, b = make_image(320, 240, 0), make_image(320, 240, 2)
    p = orc.EDLineParam(minLineLen=20)
    la, lb = orc.edline_detect(a, p, True), orc.edline_detect(b, p, True)
    r2c = orc.line_matching(a, b, la, lb)
    m = re.search(r"ok=(\d) lines_a=(\d+) lines_b=(\d+) digest_a=(\d+) digest_b=(\d+) matched=(\d+) match_digest=(\d+)", r.stdout)
    assert m, r.stdout
    assert int(m.group(1)) == 1 and int(m.group(2)) == len(la) and int(m.group(3)) == len(lb)
    assert int(m.group(4)) == line_digest(la) and int(m.group(5)) == line_digest(lb)
    assert int(m.group(6)) == int((r2c >= 0).sum()) and int(m.group(7)) == fnv(r2c + 7)
    assert len(la) >= 3 and (r2c >= 0).sum() >= 1
    m2 = re.search(r"unsmoothed_lines=(\d+) unsmoothed_digest=(\d+) empty_returns=(\d)", r.stdout)
    lu = orc.edline_detect(a, p, False)
    assert int(m2.group(1)) == len(lu) and int(m2.group(2)) == line_digest(lu) and m2.group(3) == "0"
    # vanishing_point_detection::run_vanishing_point_detection on both frames' lines, one object (frame_count 0, 1)
    for frame, ln in enumerate((la, lb)):
        mv = re.search(r"vp frame=%d n=(\d+) vps_digest=(\d+) ids_digest=(\d+) status=(-?\d+)" % frame, r.stdout)
        assert mv, r.stdout
        vps, idx, d = orc.vp_detect(ln, None, 230.0, 160.0, 120.0, 1700000123, frame, math_mode=1, details=True)
        assert int(mv.group(1)) == len(ln) and int(mv.group(2)) == fnv(vps.view(np.uint8).reshape(-1))
        assert int(mv.group(3)) == fnv(idx) and int(mv.group(4)) == (d["flags"] & 1)
    # the C++ batch driver of the fused readImage pipeline: whole run == two shards == the oracle chain
    mr = re.search(r"readimage lines=(\d+) matched=(\d+) labelled=(\d+) digest=(\d+) sharded_equal=(\d)", r.stdout)
    assert mr and mr.group(5) == "1", r.stdout
    seq = [make_image(320, 240, 2 * i) for i in range(5)]
    ls = [orc.edline_detect(f, p, True) for f in seq]
    nmatch = sum(int((orc.line_matching(seq[i - 1], seq[i], ls[i - 1], ls[i]) >= 0).sum()) for i in range(1, 5))
    nlab = sum(int((orc.vp_detect(ls[i], None, 230.0, 160.0, 120.0, 900 + i, i, math_mode=1)[1] != 3).sum()) for i in range(5))
    assert int(mr.group(1)) == sum(len(x) for x in ls) and int(mr.group(2)) == nmatch and int(mr.group(3)) == nlab


def test_tracker_facade_compiles_and_fails_loudly_without_gpu(vpl, tmp_path):
    """compat/linefeature_tracker_b200.hpp (the reference's LineFeatureTracker over the C ABI) compiles without OpenCV;
    without a device it refuses to run (no CPU path)."""
    exe = build_facade(tmp_path, "test_tracker")
    assert subprocess.run([exe, "--compile-only"]).returncode == 0
    if vpl.capi.device_count() == 0:
        r = subprocess.run([exe, "in.bin", "out.bin"], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU path" in r.stdout
