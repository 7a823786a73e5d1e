"""SURVEY 8a-R1 on the device: the reference-named LineFeatureTracker facade (compat/linefeature_tracker_b200.hpp,
compiled C++ over the C ABI) run on the bundled 15-frame EuRoC sequence as the tracker node runs the reference's
class, against tests/golden/ref_tracker.npz -- what the reference's OWN line_feature_tracker.cpp produced for the same
sequence and seeds.  Bar: every Line byte, every id, the order of the lines, the track counts (except the entry the
reference reads out of bounds); vanishing-point vectors within 1e-15 of the reference's libm result and bit-equal to
the oracle in the device's arithmetic."""
import os
import subprocess

import numpy as np
import pytest

from test_oracle_preproc import euroc_maps
from test_oracle_tracker import run_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vplines-slam_b200")


def run_facade(tmp_path, frames, seeds, K, cfg, cfg_f):
    exe = str(tmp_path / "test_tracker")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_tracker.cpp"),
                           "-L", PKG, "-lvplines_b200", f"-Wl,-rpath,{PKG}"])
    mapx, mapy = euroc_maps()
    n, h, w = frames.shape
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(np.array([n, w, h], np.int32).tobytes()); f.write(np.asarray(K, np.float32).tobytes())
        f.write(np.asarray(cfg, np.int32).tobytes()); f.write(np.asarray(cfg_f, np.float32).tobytes())
        f.write(mapx.tobytes()); f.write(mapy.tobytes()); f.write(np.asarray(seeds, np.uint32).tobytes())
        f.write(np.ascontiguousarray(frames).tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(fout, "rb").read()
    out, off = [], 0
    line_dt = np.dtype([("endpoint", "<f4", 4), ("equation", "<f8", 3), ("center", "<f4", 2), ("length", "<f4"), ("pad", "<f4")])
    for _ in range(n):
        nl, nv, nt, ex, st = np.frombuffer(raw, np.int32, 5, off); off += 20
        lines = np.frombuffer(raw, line_dt, nl, off); off += 56 * nl
        ids = np.frombuffer(raw, np.int32, nl, off); off += 4 * nl
        vps = np.frombuffer(raw, np.float64, 4 * nv, off).reshape(nv, 4); off += 32 * nv
        t = np.frombuffer(raw, np.int32, nt, off); off += 4 * nt
        out.append(dict(lines=lines, ids=ids, vps=vps, t_cnt=t, lines_exit=bool(ex), status=int(st)))
    assert off == len(raw)
    return out


def test_tracker_facade_equals_reference(vpl, orc, mh04, tmp_path):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_tracker.npz"))
    for s0 in (int(s) for s in gold["seeds"]):
        seeds = [s0 + i for i in range(len(mh04))]
        got = run_facade(tmp_path, mh04, seeds, gold["K"], gold["cfg"], gold["cfg_f"])
        t, want = run_oracle(orc, mh04, gold, s0, 1)
        oob = {n[1] for n in t.notes if n[0] == "t_cnt_out_of_range"}
        for i, g in enumerate(got):
            p = "s%d_f%02d_" % (s0, i)
            assert g["lines_exit"] and g["status"] == 0
            # the reference's own output: lines (every byte but the padding), order, ids
            gl = gold[p + "lines"]
            assert len(g["lines"]) == len(gl), p
            for k in ("endpoint", "equation", "center", "length"):
                assert g["lines"][k].tobytes() == gl[k].tobytes(), (p, k)
            assert list(g["ids"]) == list(gold[p + "ids"]), p
            assert g["vps"].shape == gold[p + "vps"].shape and np.allclose(g["vps"], gold[p + "vps"], rtol=0, atol=1e-15), p
            bad = [j for j in range(len(g["t_cnt"])) if g["t_cnt"][j] != gold[p + "t_cnt"][j]]
            assert len(g["t_cnt"]) == len(gold[p + "t_cnt"]) and set(bad) <= oob, (p, bad)
            # the oracle in the device's arithmetic: vanishing-point vectors bit for bit
            assert g["vps"].tobytes() == np.asarray(want[i]["vps"], np.float64).reshape(-1, 4).tobytes(), p
