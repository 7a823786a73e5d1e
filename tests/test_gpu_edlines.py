"""GPU parity for SURVEY 8f-1: the CUDA EDLines detector (vplines-slam_b200/csrc/edlines.cu, through
the C ABI vpl_edlines_*) against (1) the golden vectors produced by the reference's own
edline_detector.cpp (tests/golden/ref_edlines.npz) and (2) the CPU oracle on seeded inputs.
Bar: bit-exact -- edge-chain pixels, chain starts and every byte of every Line record."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_edlines", os.path.join(HERE, "golden", "make_golden_edlines.py"))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)
CASES = mk.cases()


@pytest.fixture(scope="module")
def ctx(vpl):
    c = vpl.Context(max_width=752, max_height=480, max_octaves=1, max_lines=1024, max_batch=32, num_slots=2,
                    profile=True)
    yield c
    c.close()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_edlines.npz"))


def to_capi(vpl, p):
    return vpl.capi.EDLineParam(p.ksize, p.sigma, p.gradientThreshold, p.anchorThreshold, p.scanIntervals,
                                p.minLineLen, p.lineFitErrThreshold)


@pytest.mark.parametrize("name", sorted(CASES))
def test_edlines_reference_golden(ctx, vpl, gold, name):
    img, p, sm = CASES[name]
    ctx.edlines_configure(to_capi(vpl, p))
    lines, status = ctx.edlines_detect_batch([img], smoothed=sm, with_status=True)
    xy, sid = ctx.edge_chains(0, img.shape[1], img.shape[0])
    assert np.array_equal(xy, gold[name + "_xy"])
    assert np.array_equal(sid, gold[name + "_sid"])
    assert lines[0].tobytes() == gold[name + "_lines"].tobytes()
    assert status[0] == (1 if len(gold[name + "_xy"]) else -1)


def test_edlines_batch_vs_oracle_real_frames(ctx, vpl, orc, mh04):
    p = orc.EDLineParam()
    ctx.edlines_configure(to_capi(vpl, p))
    got = ctx.edlines_detect_batch(mh04, smoothed=True)
    total = 0
    for f in range(len(mh04)):
        exp = orc.edline_detect(mh04[f], p, True)
        assert got[f].tobytes() == exp.tobytes(), f"frame {f}"
        total += len(exp)
    assert total > 1500


def test_edlines_batch_vs_oracle_synthetic_unsmoothed(ctx, vpl, orc, synth):
    frames = synth.sequence(24, w=376, h=240, seed=31, n_quads=12, n_strokes=20)
    p = orc.EDLineParam(minLineLen=20, lineFitErrThreshold=1.4)
    ctx.edlines_configure(to_capi(vpl, p))
    got = ctx.edlines_detect_batch(frames, smoothed=False)
    for f in range(len(frames)):
        assert got[f].tobytes() == orc.edline_detect(frames[f], p, False).tobytes(), f"frame {f}"


def test_edlines_pipelined_slots_and_resident(ctx, vpl, orc, mh04):
    p = orc.EDLineParam()
    ctx.edlines_configure(to_capi(vpl, p))
    cap = 512
    bufs = []
    for slot, sl in enumerate((slice(0, 8), slice(8, 15))):
        n = ctx.edlines_submit(slot, mh04[sl], smoothed=True)
        bufs.append((np.zeros((n, cap), vpl.capi.LINE_DTYPE), np.zeros(n, np.int32), np.zeros(n, np.int32)))
    for slot, sl in enumerate((slice(0, 8), slice(8, 15))):
        lines, counts, status = bufs[slot]
        ctx.edlines_collect_into(slot, lines, counts, cap, status)
        for i, f in enumerate(range(sl.start, sl.stop)):
            assert lines[i, :counts[i]].tobytes() == orc.edline_detect(mh04[f], p, True).tobytes()
            assert status[i] == 1
    ctx.edlines_run_resident(0)  # re-run on the resident frames: must not fault
    ctx.sync()
    t = ctx.stage_times()
    assert t["ed_walk"][1] >= 1 and t["ed_fit"][1] >= 1


def test_edlines_failure_cases(ctx, vpl, orc):
    ctx.edlines_configure(to_capi(vpl, orc.EDLineParam()))
    flat = np.full((64, 80), 9, np.uint8)
    lines, status = ctx.edlines_detect_batch([flat], smoothed=True, with_status=True)
    assert len(lines[0]) == 0 and status[0] == -1            # "lines not found", edline_detector.cpp:667
    # more anchors than W*H/5 (edline_detector.cpp:166-169): the reference writes out of bounds
    # and returns -1; the device reports the frame as failed, and other frames are unaffected
    rng = np.random.default_rng(1)
    noise = rng.integers(0, 256, (96, 128), dtype=np.uint8)
    step = np.full((96, 128), 50, np.uint8); step[:, 61] = 125; step[:, 62:] = 200
    p = orc.EDLineParam(scanIntervals=1, anchorThreshold=0, gradientThreshold=0, minLineLen=5)
    ctx.edlines_configure(to_capi(vpl, p))
    lines, status = ctx.edlines_detect_batch([noise, step], smoothed=True, with_status=True)
    assert len(lines[0]) == 0 and status[0] == -1
    assert lines[1].tobytes() == orc.edline_detect(step, p, True).tobytes()
    # unsupported blur parameters are refused, not approximated
    ctx.edlines_configure(to_capi(vpl, orc.EDLineParam(ksize=7, sigma=1.5)))
    with pytest.raises(vpl.VplError):
        ctx.edlines_detect_batch([flat], smoothed=False)
    with pytest.raises(vpl.VplError):
        ctx.edlines_detect_batch([np.zeros((500, 800), np.uint8)], smoothed=True)  # larger than the context


def test_edlines_line_capacity_error(vpl, orc, mh04):
    with vpl.Context(max_width=752, max_height=480, max_lines=16, max_batch=2) as c:
        c.edlines_configure(to_capi(vpl, orc.EDLineParam()))
        with pytest.raises(vpl.VplError) as e:
            c.edlines_detect_batch(mh04[:1], smoothed=True)
        assert e.value.code == vpl.capi.VPL_E_CAPACITY
