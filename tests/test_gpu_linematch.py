"""GPU parity for SURVEY 8f-2: the CUDA line matcher (vplines-slam_b200/csrc/linematch.cu, through the
C ABI vpl_linematch_* / vpl_linefront_*) against (1) the golden vectors produced by the reference's
own line_matching.cpp + lk_tracker_invoker_2d.cpp (tests/golden/ref_linematch.npz) and (2) the CPU
oracle on seeded inputs.  Bar: bit-exact -- anchors, tracked positions (float32), status, error,
closest-line labels and the match vector."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_linematch", os.path.join(HERE, "golden", "make_golden_linematch.py"))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)
CASES = mk.cases()


@pytest.fixture(scope="module")
def ctx(vpl):
    c = vpl.Context(max_width=752, max_height=480, max_octaves=1, max_lines=512, max_batch=32, num_slots=2,
                    profile=True, lsd_path=False)
    yield c
    c.close()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_linematch.npz"))


def same_bits(a, b):
    """bit-identical; NaNs (which x86 and the GPU encode differently) must sit at the same places"""
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        return np.array_equal(na, nb) and a[~na].tobytes() == b[~nb].tobytes()
    return a.tobytes() == b.tobytes()


def as_capi(vpl, lines):
    return np.ascontiguousarray(lines).view(vpl.capi.LINE_DTYPE).reshape(-1)


@pytest.mark.parametrize("name", sorted(CASES))
def test_linematch_reference_golden(ctx, vpl, gold, name):
    a, b, p, illum, topo = CASES[name]
    ctx.linematch_configure(vpl.capi.LineMatchParam(illumination_adapt=int(illum), topological_filter=int(topo)))
    la, lb = as_capi(vpl, gold[name + "_lines_ref"]), as_capi(vpl, gold[name + "_lines_cur"])
    r2c = ctx.linematch_batch([a], [b], [la], [lb])[0]
    d = ctx.linematch_points(0)
    for k in ("kps_ref", "kps_cur", "status", "err", "kp2line"):
        assert same_bits(d[k], gold[name + "_" + k]), k
    assert np.array_equal(r2c, gold[name + "_ref_to_cur"])


def test_linematch_batch_of_pairs_vs_oracle(ctx, vpl, orc, mh04):
    ctx.linematch_configure(vpl.capi.LineMatchParam())
    p = orc.EDLineParam()
    lines = [orc.edline_detect(f, p, True) for f in mh04]
    idx = list(range(0, 14))
    got = ctx.linematch_batch([mh04[i] for i in idx], [mh04[i + 1] for i in idx], [as_capi(vpl, lines[i]) for i in idx],
                              [as_capi(vpl, lines[i + 1]) for i in idx])
    total = 0
    for k, i in enumerate(idx):
        exp = orc.line_matching(mh04[i], mh04[i + 1], lines[i], lines[i + 1])
        assert np.array_equal(got[k], exp), f"pair {i}"
        total += int((exp >= 0).sum())
    assert total > 800


def test_linefront_fused_sequence_vs_oracle(ctx, vpl, orc, mh04):
    """EDLines + matching fused on a frame sequence == oracle EDLines + oracle matching."""
    ctx.edlines_configure(vpl.capi.EDLineParam())
    ctx.linematch_configure(vpl.capi.LineMatchParam())
    lines, p2c = ctx.linefront_batch(mh04, smoothed=True)
    p = orc.EDLineParam()
    exp_lines = [orc.edline_detect(f, p, True) for f in mh04]
    for f in range(len(mh04)):
        assert lines[f].tobytes() == exp_lines[f].tobytes(), f"lines {f}"
        if f:
            exp = orc.line_matching(mh04[f - 1], mh04[f], exp_lines[f - 1], exp_lines[f])
            assert np.array_equal(p2c[f], exp), f"match {f}"
    t = ctx.stage_times()
    assert t["lm_track"][1] >= 1 and t["lm_vote"][1] >= 1


def test_linefront_pipelined_slots(ctx, vpl, orc, mh04):
    ctx.edlines_configure(vpl.capi.EDLineParam())
    ctx.linematch_configure(vpl.capi.LineMatchParam())
    cap = 512
    chunks = (slice(0, 8), slice(7, 15))  # one-frame overlap: every consecutive pair is matched once
    bufs = []
    for slot, sl in enumerate(chunks):
        n = ctx.linefront_submit(slot, mh04[sl], smoothed=True)
        bufs.append((np.zeros((n, cap), vpl.capi.LINE_DTYPE), np.zeros(n, np.int32), np.full((n, cap), -1, np.int32)))
    p = orc.EDLineParam()
    exp_lines = [orc.edline_detect(f, p, True) for f in mh04]
    for slot, sl in enumerate(chunks):
        lines, counts, p2c = bufs[slot]
        ctx.linefront_collect_into(slot, lines, counts, cap, p2c)
        for i, f in enumerate(range(sl.start, sl.stop)):
            assert lines[i, :counts[i]].tobytes() == exp_lines[f].tobytes()
            if i:
                exp = orc.line_matching(mh04[f - 1], mh04[f], exp_lines[f - 1], exp_lines[f])
                assert np.array_equal(p2c[i, :counts[i - 1]], exp)
    ctx.linefront_run_resident(0)
    ctx.sync()


def test_linematch_edge_cases(ctx, vpl, orc, mh04, synth):
    ctx.linematch_configure(vpl.capi.LineMatchParam())
    la = as_capi(vpl, orc.edline_detect(mh04[0]))
    # an empty side: Matching returns false -> nothing matched
    r = ctx.linematch_batch([mh04[0], mh04[0]], [mh04[1], mh04[1]], [la, la[:0]], [la[:0], la])
    assert len(r[0]) == len(la) and (r[0] == -1).all() and len(r[1]) == 0
    # small image: the pyramid stops early (a 20x13 level would not exceed the 13-px window)
    a = np.ascontiguousarray(mh04[6][200:300, 300:460]); b = np.ascontiguousarray(mh04[7][200:300, 300:460])
    p = orc.EDLineParam(minLineLen=15)
    l1, l2 = orc.edline_detect(a, p, True), orc.edline_detect(b, p, True)
    got = ctx.linematch_batch([a], [b], [as_capi(vpl, l1)], [as_capi(vpl, l2)])[0]
    assert np.array_equal(got, orc.line_matching(a, b, l1, l2))
    # anchors leaving the image (lines touching the border, large shift): status / error paths
    s = synth.sequence(2, w=320, h=200, seed=5, n_quads=10, n_strokes=16)
    shifted = np.roll(s[0], 37, axis=1)
    pp = orc.EDLineParam(minLineLen=18)
    l1, l2 = orc.edline_detect(s[0], pp, True), orc.edline_detect(shifted, pp, True)
    got = ctx.linematch_batch([s[0]], [shifted], [as_capi(vpl, l1)], [as_capi(vpl, l2)])[0]
    exp, d = orc.line_matching(s[0], shifted, l1, l2, details=True)
    dg = ctx.linematch_points(0)
    assert np.array_equal(got, exp) and same_bits(dg["kps_cur"], d["kps_cur"]) and same_bits(dg["status"], d["status"])
    # anchor capacity
    ctx.linematch_configure(vpl.capi.LineMatchParam(max_anchors=64))
    with pytest.raises(vpl.VplError) as e:
        ctx.linematch_batch([mh04[0]], [mh04[1]], [la], [la])
    assert e.value.code == vpl.capi.VPL_E_CAPACITY


def test_lsd_entry_points_refused_without_lsd_path(ctx, vpl, mh04):
    with pytest.raises(vpl.VplError):
        ctx.lsd_detect_batch(mh04[:1])


@pytest.mark.parametrize("shape,scan,minlen", [((720, 1280), 2, 35), ((207, 331), 1, 15), ((40, 52), 2, 10)])
def test_linefront_other_geometries(vpl, orc, synth, mh04, shape, scan, minlen):
    """D455-shaped 1280x720 frames, an odd size scanned at every pixel, and an image so small that the
    pyramid has two levels (26x20 is the last one above the 13-px window): fused EDLines + matching ==
    oracle, bit for bit."""
    h, w = shape
    if h < 100:
        frames = np.ascontiguousarray(mh04[5:8, 150:150 + h, 350:350 + w])
    else:
        big = synth.sequence(3, w=max(w, 376), h=max(h, 240), seed=900 + h, n_quads=14, n_strokes=24)
        frames = np.ascontiguousarray(big[:, :h, :w])
    pe = orc.EDLineParam(minLineLen=minlen, scanIntervals=scan)
    with vpl.Context(max_width=w, max_height=h, max_lines=1024, max_batch=4, num_slots=1, lsd_path=False) as c:
        c.edlines_configure(vpl.capi.EDLineParam(minLineLen=minlen, scanIntervals=scan))
        c.linematch_configure(vpl.capi.LineMatchParam(max_anchors=16384))
        lines, p2c = c.linefront_batch(frames, smoothed=False)
    exp_lines = [orc.edline_detect(f, pe, False) for f in frames]
    matched = 0
    for f in range(len(frames)):
        assert lines[f].tobytes() == exp_lines[f].tobytes(), f"lines {f}"
        if f:
            exp = orc.line_matching(frames[f - 1], frames[f], exp_lines[f - 1], exp_lines[f])
            if exp is None:
                assert (p2c[f] == -1).all()
            else:
                assert np.array_equal(p2c[f], exp), f"match {f}"
                matched += int((exp >= 0).sum())
    assert sum(len(l) for l in exp_lines) > 0


def test_linematch_non_default_parameters(ctx, vpl, orc, mh04):
    """step 6, looser thresholds, 2 pyramid levels, 10 iterations: the parameter plumbing."""
    kw = dict(step=6, closest_line_threshold=1.0, line_matching_ratio=0.3, line_distance_error_ratio=2.0,
              klt_error_threshold=25.0, max_level=1, max_count=10, epsilon=0.01, min_eig=1e-3,
              topo_distance_threshold=10.0, topo_length_ratio=0.3, topo_violation_ratio=0.1)
    ctx.linematch_configure(vpl.capi.LineMatchParam(max_anchors=8192, **kw))
    p = orc.EDLineParam()
    la, lb = orc.edline_detect(mh04[4], p, True), orc.edline_detect(mh04[5], p, True)
    got = ctx.linematch_batch([mh04[4]], [mh04[5]], [as_capi(vpl, la)], [as_capi(vpl, lb)])[0]
    d = ctx.linematch_points(0)
    exp, de = orc.line_matching(mh04[4], mh04[5], la, lb, param=orc.LineMatchParam(**kw), details=True)
    for k in ("kps_ref", "kps_cur", "status", "err", "kp2line"):
        assert same_bits(d[k], de[k]), k
    assert np.array_equal(got, exp) and (exp >= 0).sum() > 20
