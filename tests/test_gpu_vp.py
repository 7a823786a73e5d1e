"""GPU parity for SURVEY 8f-4: the CUDA vanishing-point stage (vplines-slam_b200/csrc/vp.cu, through
the C ABI vpl_vp_*) against (1) the golden vectors produced by the reference's own
vanishing_point_detection.cpp (tests/golden/ref_vp.npz) and (2) the CPU oracle on seeded inputs.
Bar: against the oracle in its device arithmetic (math_mode 1: the shared deterministic
atan/acos/sincos) everything is bit-exact -- the smoothed sphere grid (32 400 doubles), the random
line pairs, the winning hypothesis, the nine doubles of the vanishing points, every line label and
the out-of-range flag; against the reference's own output (computed with libm) labels are identical
and the vanishing points agree to 1e-15 (tolerance stated: unit vectors, i.e. <= 5 ulp)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
EUROC = (461.6, 363.0, 248.1)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_vp.npz"))


NAMES = sorted(n[:-6] for n in np.load(os.path.join(HERE, "golden", "ref_vp.npz")).files
               if n.endswith("_lines") and not n.endswith("_all_lines"))


@pytest.fixture(scope="module")
def ctx(vpl):
    c = vpl.Context(max_width=752, max_height=480, max_octaves=1, max_lines=2048, max_batch=64, num_slots=2,
                    profile=True, lsd_path=False)
    yield c
    c.close()


def as_capi(vpl, lines):
    return np.ascontiguousarray(lines).view(vpl.capi.LINE_DTYPE).reshape(-1)


def check_frame(orc, dev_vps, dev_idx, dev_status, dbg, ln, al, f, cx, cy, seed, fc, scores=None):
    vps, idx, d = orc.vp_detect(ln, al, f, cx, cy, seed, fc, math_mode=1, details=True)
    assert dev_status == (d["flags"] & 1)
    assert np.array_equal(dev_idx, idx)
    assert dev_vps.tobytes() == vps.tobytes()
    if dbg is not None:
        assert dbg["best_idx"] == d["best_idx"]
        assert np.array_equal(dbg["pairs"], d["pairs"])
        assert dbg["grid"].tobytes() == d["grid"].tobytes()
    if scores is not None:  # every one of the 37 800 hypotheses read the same three cells
        assert scores.tobytes() == d["scores"].tobytes()
    return d


@pytest.mark.parametrize("name", NAMES)
def test_vp_reference_golden(ctx, vpl, orc, gold, name):
    f, cx, cy = (float(v) for v in gold[name + "_cam"])
    seed, fc = (int(v) for v in gold[name + "_seed"])
    ln, al = gold[name + "_lines"], gold[name + "_all_lines"]
    ctx.vp_configure(f, cx, cy)
    same = ln.tobytes() == al.tobytes()
    vps, idx, st = ctx.vp_detect_batch([as_capi(vpl, ln)], [seed], None if same else [as_capi(vpl, al)], frame_count0=fc)
    # the reference's own output: same labels, vanishing points to 1e-15
    assert st[0] == 0
    assert np.array_equal(idx[0], gold[name + "_vp_idx"])
    assert np.abs(vps[0] - gold[name + "_vps"]).max() <= 1e-15
    # the oracle in the device's arithmetic: bit-exact, stage by stage
    check_frame(orc, vps[0], idx[0], st[0], ctx.vp_debug(0), ln, al, f, cx, cy, seed, fc, scores=ctx.vp_scores(0))


def test_vp_batch_vs_oracle_many_seeds(ctx, vpl, orc, mh04):
    """60 (frame, seed) combinations in one batch, frame_count running 0, 1, 2, ... inside the batch; includes
    frames on which the reference reads lx[] out of range (status 1: repaired semantics of oracle/orc_vp.c)."""
    ctx.vp_configure(*EUROC)
    sets, seeds = [], []
    for k in range(15):
        ln = orc.edline_detect(mh04[k])
        for s in range(4):
            sets.append(ln); seeds.append(1700000000 + 97 * k + s)
    vps, idx, st = ctx.vp_detect_batch([as_capi(vpl, l) for l in sets], seeds, frame_count0=0)
    n_flag = 0
    for i, (ln, seed) in enumerate(zip(sets, seeds)):
        d = check_frame(orc, vps[i], idx[i], st[i], ctx.vp_debug(i) if i % 7 == 0 else None, ln, ln, *EUROC, seed, i,
                        scores=ctx.vp_scores(i) if i % 3 == 0 else None)
        n_flag += d["flags"]
    assert 0 < n_flag < len(sets)          # both kinds of frame were exercised
    assert (np.concatenate(idx) != 3).sum() > 1000


def test_vp_subset_lines_and_line_vps(ctx, vpl, orc, mh04):
    """`lines` = the near-vertical subset, all_lines = everything (line_feature_tracker.cpp:241); per-line Vector4d."""
    ctx.vp_configure(*EUROC)
    subs, alls, seeds = [], [], []
    for k in (1, 6, 11):
        ln = orc.edline_detect(mh04[k])
        e = ln["endpoint"]
        vert = np.abs(e[:, 0] - e[:, 2]) < 0.35 * np.abs(e[:, 1] - e[:, 3])
        subs.append(ln[vert]); alls.append(ln); seeds.append(40 + k)
    vps, idx, st, lv = ctx.vp_detect_batch([as_capi(vpl, l) for l in subs], seeds, [as_capi(vpl, l) for l in alls],
                                           frame_count0=4, with_line_vps=True)
    for i in range(3):
        check_frame(orc, vps[i], idx[i], st[i], ctx.vp_debug(i), subs[i], alls[i], *EUROC, seeds[i], 4 + i)
        for j, lab in enumerate(idx[i]):  # line_feature_tracker.cpp:246-262
            want = np.zeros(4) if lab == 3 else np.array([*vps[i][lab], vps[i][lab][2] / vps[i][lab][2]])
            assert lv[i][j].tobytes() == want.tobytes()


def test_vp_many_lines(ctx, vpl, orc):
    """1500 synthetic segments towards three vanishing points (1.1 M pairs per frame; cells of the vote
    receive thousands of additions in pair order), 1280x720 intrinsics."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mkvp", os.path.join(HERE, "golden", "make_golden_vp.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    ctx.vp_configure(640.0, 640.0, 360.0)
    sets = [mk.manhattan_lines(1500, seed=21), mk.manhattan_lines(700, seed=22)]
    seeds = [11, 1700001234]
    vps, idx, st = ctx.vp_detect_batch([as_capi(vpl, l) for l in sets], seeds, frame_count0=0)
    for i in range(2):
        check_frame(orc, vps[i], idx[i], st[i], ctx.vp_debug(i), sets[i], sets[i], 640.0, 640.0, 360.0, seeds[i], i)
    assert (idx[0] != 3).sum() > 500


def test_vp_degenerate_frames(ctx, vpl, orc, mh04):
    """0 / 1 lines -> status -1, nothing labelled, vps zero (readImage's "no vp lines" branch); 2 and 3 lines run;
    parallel lines only (every intersection has z == 0) -> status -2 where the reference would loop for ever."""
    ctx.vp_configure(*EUROC)
    ln = orc.edline_detect(mh04[3])
    par = np.zeros(3, orc.LINE_DTYPE)
    for i in range(3):
        par["endpoint"][i] = (10, 20 + 30 * i, 210, 20 + 30 * i)
    sets = [ln[:0], ln[:1], ln[:2], ln[:3], par, ln]
    seeds = [5, 6, 7, 8, 9, 10]
    vps, idx, st = ctx.vp_detect_batch([as_capi(vpl, l) for l in sets], seeds, frame_count0=0)
    assert list(st[:2]) == [-1, -1] and (vps[:2] == 0).all() and len(idx[0]) == 0 and list(idx[1]) == [3]
    assert st[4] == -2 and (vps[4] == 0).all() and list(idx[4]) == [3, 3, 3]
    with pytest.raises(ValueError):
        orc.vp_detect(par, seed=9)
    for i in (2, 3, 5):
        check_frame(orc, vps[i], idx[i], st[i], None, sets[i], sets[i], *EUROC, seeds[i], i)


def test_vp_submit_collect_two_slots(ctx, vpl, orc, mh04):
    ctx.vp_configure(*EUROC)
    L = vpl.capi.LINE_DTYPE
    sets = [as_capi(vpl, orc.edline_detect(mh04[k])) for k in range(8)]
    cap = max(len(s) for s in sets)
    arr = np.zeros((8, cap), L); cnt = np.array([len(s) for s in sets], np.int32)
    for i, s in enumerate(sets):
        arr[i, :cnt[i]] = s
    seeds = np.arange(900, 908, dtype=np.uint32)
    ref_vps, ref_idx, ref_st = ctx.vp_detect_batch(sets, seeds, frame_count0=0)
    ctx.vp_submit(0, arr[:4], cnt[:4], seeds[:4], 0)
    ctx.vp_submit(1, np.ascontiguousarray(arr[4:]), cnt[4:], seeds[4:], 4)
    for slot, lo in ((0, 0), (1, 4)):
        vps = np.zeros((4, 3, 3)); idx = np.full((4, cap), -1, np.int32); st = np.zeros(4, np.int32)
        ctx.vp_collect_into(slot, cap, vps, idx, st)
        assert vps.tobytes() == ref_vps[lo:lo + 4].tobytes() and np.array_equal(st, ref_st[lo:lo + 4])
        for i in range(4):
            assert np.array_equal(idx[i, :cnt[lo + i]], ref_idx[lo + i])
    t = ctx.stage_times()
    assert t["vp_vote"][1] >= 2 and t["vp_score"][1] >= 1 and t["vp_classify"][1] >= 1


def test_vp_errors(ctx, vpl):
    c2 = vpl.Context(max_width=64, max_height=64, max_lines=16, max_batch=2, lsd_path=False)
    with pytest.raises(vpl.capi.VplError):
        c2.vp_detect_batch([np.zeros(3, vpl.capi.LINE_DTYPE)], [1])      # not configured
    c2.vp_configure(100.0, 32.0, 32.0)
    with pytest.raises(vpl.capi.VplError):
        c2.vp_detect_batch([np.zeros(17, vpl.capi.LINE_DTYPE)], [1])     # more lines than max_lines
    with pytest.raises(vpl.capi.VplError):
        c2.vp_detect_batch([np.zeros(3, vpl.capi.LINE_DTYPE)] * 3, [1, 2, 3])  # more frames than max_batch
    c2.close()


def test_vp_pointcloud_packing(ctx, vpl, orc, mh04):
    """vpl_vp_pack_cloud against the oracle restatement of img_callback's packing loop: bit-exact floats."""
    ctx.vp_configure(*EUROC)
    sets = [orc.edline_detect(mh04[k]) for k in (2, 7)] + [orc.edline_detect(mh04[4])[:1]]
    ids = [np.arange(len(s), dtype=np.int32) * 3 + 11 * (i + 1) for i, s in enumerate(sets)]
    vps, idx, st, lv = ctx.vp_detect_batch([as_capi(vpl, l) for l in sets], [31, 32, 33], frame_count0=2, with_line_vps=True)
    fx, fy, cx, cy = 461.6, 460.3, 363.0, 248.1
    for ncam, cam in ((1, 0), (2, 1)):
        got = ctx.vp_pack_cloud(ids, fx, fy, cx, cy, ncam, cam)
        for i, s in enumerate(sets):
            want = orc.line_cloud(s, ids[i], lv[i], fx, fy, cx, cy, ncam, cam)
            for k in want:
                assert got[i][k].tobytes() == want[k].tobytes(), (i, k)
            assert (got[i]["points"][:, 2] == 1).all()
    # every line carries the Vector4d of line number `cam` (the reference indexes the per-line list with the camera index)
    g = ctx.vp_pack_cloud(ids, fx, fy, cx, cy, 1, 0)[0]
    assert (g["vp_x"] == np.float32(lv[0][0][0])).all()


def _fuzz_sets(orc, rng, n_sets):
    """Line sets with the degenerate shapes a tracker can hand over: duplicates and collinear segments (intersections
    with z == 0, redrawn pairs), zero-length segments (NaN directions), axis-parallel bundles (no two lines intersect),
    far-away coordinates, very few lines."""
    sets = []
    for s in range(n_sets):
        n = int(rng.integers(2, 60))
        ln = np.zeros(n, orc.LINE_DTYPE)
        e = rng.uniform(0, 700, (n, 4)).astype(np.float32)
        kind = s % 6
        if kind == 1:                       # half of the segments duplicated
            e[n // 2:] = e[:n - n // 2]
        elif kind == 2:                     # collinear runs on a few carrier lines
            for i in range(n):
                a, b = rng.uniform(0, 600, 2)
                k = i % 3
                e[i] = (a, 50 + 100 * k + 0.25 * a, b, 50 + 100 * k + 0.25 * b)
        elif kind == 3:                     # some zero-length segments
            e[::4, 2:] = e[::4, :2]
        elif kind == 4:                     # integer endpoints, many exactly parallel
            e = np.round(e / 50) * 50
            e[:, 3] = e[:, 1] + (e[:, 2] - e[:, 0])
        elif kind == 5:                     # far away / tiny
            e[: n // 3] *= 1e4
            e[n // 3: 2 * n // 3] *= 1e-3
        ln["endpoint"] = e
        ln["center"] = (e[:, :2] + e[:, 2:]) / 2
        ln["length"] = np.hypot(e[:, 2] - e[:, 0], e[:, 3] - e[:, 1])
        sets.append(ln)
    return sets


def test_vp_fuzz_degenerate_line_sets(ctx, vpl, orc):
    """120 degenerate line sets, device vs oracle (device arithmetic): status, labels, vanishing points and the whole
    grid bit for bit -- NaNs included (a zero-length segment has no direction: its three angles are NaN, label 3)."""
    ctx.vp_configure(*EUROC)
    rng = np.random.default_rng(2024)
    sets = _fuzz_sets(orc, rng, 120)
    seeds = rng.integers(1, 2**32 - 1, len(sets), dtype=np.uint64).astype(np.uint32)
    n_err = n_ok = 0
    for lo in range(0, len(sets), 60):
        chunk = sets[lo:lo + 60]
        vps, idx, st = ctx.vp_detect_batch([as_capi(vpl, l) for l in chunk], seeds[lo:lo + 60], frame_count0=lo)
        for i, ln in enumerate(chunk):
            try:
                ev, ei, d = orc.vp_detect(ln, None, *EUROC, int(seeds[lo + i]), lo + i, math_mode=1, details=True)
            except ValueError:              # the oracle gave up (-2: no two lines intersect)
                assert st[i] == -2 and (idx[i] == 3).all() and (vps[i] == 0).all()
                n_err += 1
                continue
            assert st[i] == (d["flags"] & 1), (lo + i, st[i], d["flags"])
            assert np.array_equal(idx[i], ei), lo + i
            a, b = vps[i], ev
            assert np.array_equal(np.isnan(a), np.isnan(b)) and a[~np.isnan(a)].tobytes() == b[~np.isnan(b)].tobytes(), lo + i
            if i % 5 == 0:
                g = ctx.vp_debug(i)["grid"]
                assert np.array_equal(np.isnan(g), np.isnan(d["grid"])) and g[~np.isnan(g)].tobytes() == d["grid"][~np.isnan(g)].tobytes()
            n_ok += 1
    assert n_ok >= 90


def test_readimage_fused_pipeline(vpl, orc, mh04):
    """vpl_readimage_*: remap + CLAHE -> EDLines -> Matching(f-1, f) -> vanishing points on each frame's own lines, one
    pass over the device, against the oracle chain stage by stage (bit-exact), with and without pre-processing; the
    resident re-run reproduces the submitted run."""
    from test_oracle_preproc import euroc_maps
    mapx, mapy = euroc_maps()
    frames = mh04[3:8]
    seeds = np.arange(77, 77 + len(frames), dtype=np.uint32)
    p = orc.EDLineParam()
    with vpl.Context(max_width=752, max_height=480, max_lines=512, max_batch=8, num_slots=2, lsd_path=False) as c:
        c.edlines_configure(vpl.capi.EDLineParam())
        c.linematch_configure(vpl.capi.LineMatchParam())
        c.vp_configure(*EUROC)
        for pre in (False, True):
            if pre:
                c.set_preprocess(mapx, mapy, clahe_clip=3.0, clahe_tiles=8)
                src = [orc.clahe(orc.remap_linear(f, mapx, mapy), 3.0, 8) for f in frames]
            else:
                src = list(frames)
            lines, p2c, vps, idx, st = c.readimage_batch(frames, seeds, smoothed=True, frame_count0=0)
            el = [orc.edline_detect(f, p, True) for f in src]
            for f in range(len(frames)):
                assert lines[f].tobytes() == el[f].tobytes(), (pre, f)
                if f:
                    assert np.array_equal(p2c[f], orc.line_matching(src[f - 1], src[f], el[f - 1], el[f])), (pre, f)
                ev, ei, d = orc.vp_detect(el[f], None, *EUROC, int(seeds[f]), f, math_mode=1, details=True)
                assert vps[f].tobytes() == ev.tobytes() and np.array_equal(idx[f], ei) and st[f] == (d["flags"] & 1), (pre, f)
            # resident re-run (from the raw frames when pre-processing is on), then the same outputs again
            c.readimage_run_resident(0)
            c.sync()
            v2, i2, s2 = c.vp_detect_batch([as_capi(vpl, l) for l in el], seeds, frame_count0=0)
            assert v2.tobytes() == vps.tobytes() and all(np.array_equal(a, b) for a, b in zip(i2, idx))


def test_slot_batch_kinds_and_cloud_after_readimage(vpl, orc, mh04):
    """Every collect requires its own kind of batch; vpl_vp_pack_cloud after a vpl_readimage_* batch packs the lines
    that batch's vanishing-point stage classified (the frame's detected lines), not the stale vpl_vp_submit buffers;
    vpl_frontend_run_resident validates k and the kind of the resident batch."""
    frames = mh04[3:6]
    seeds = np.arange(5, 5 + len(frames), dtype=np.uint32)
    fx, fy, cx, cy = 461.6, 460.3, 363.0, 248.1
    with vpl.Context(max_width=752, max_height=480, max_lines=512, max_batch=4, num_slots=2, lsd_path=False) as c:
        c.edlines_configure(vpl.capi.EDLineParam())
        c.linematch_configure(vpl.capi.LineMatchParam())
        c.vp_configure(*EUROC)
        # pack_cloud before any vanishing-point batch: refused, nothing read
        with pytest.raises(vpl.capi.VplError):
            c.vp_pack_cloud([np.zeros(3, np.int32)], fx, fy, cx, cy)
        lines, p2c, vps, idx, st = c.readimage_batch(frames, seeds, smoothed=True, frame_count0=0)
        ids = [np.arange(len(l), dtype=np.int32) * 2 + 7 for l in lines]
        got = c.vp_pack_cloud(ids, fx, fy, cx, cy, 1, 0)
        el = [orc.edline_detect(f, orc.EDLineParam(), True) for f in frames]
        v2, i2, s2, lv = c.vp_detect_batch([as_capi(vpl, l) for l in el], seeds, frame_count0=0, with_line_vps=True)
        assert v2.tobytes() == vps.tobytes()
        for i, s in enumerate(el):
            want = orc.line_cloud(s, ids[i], lv[i], fx, fy, cx, cy, 1, 0)
            for k in want:
                assert got[i][k].tobytes() == want[k].tobytes(), (i, k)
        # a readImage batch in flight is not a line-front-end / EDLines / vanishing-point batch
        n = c.readimage_submit(1, frames, seeds)
        cap = c.max_lines
        ln = np.zeros((n, cap), vpl.capi.LINE_DTYPE); cnt = np.zeros(n, np.int32); pc = np.full((n, cap), -1, np.int32)
        vv = np.zeros((n, 3, 3)); ii = np.full((n, cap), -1, np.int32)
        with pytest.raises(vpl.capi.VplError):
            c.linefront_collect_into(1, ln, cnt, cap, pc)
        with pytest.raises(vpl.capi.VplError):
            c.edlines_collect_into(1, ln, cnt, cap)
        with pytest.raises(vpl.capi.VplError):
            c.vp_collect_into(1, cap, vv, ii)
        c.readimage_collect_into(1, ln, cnt, cap, pc, vv, ii)
        assert vv.tobytes() == vps.tobytes() and all(ln[f, :cnt[f]].tobytes() == lines[f].tobytes() for f in range(n))
        # nothing is left in flight and the slot can be reused for any kind
        got_lines = c.edlines_detect_batch(frames[:1], smoothed=True)
        assert got_lines[0].tobytes() == el[0].tobytes()
    with vpl.Context(max_width=752, max_height=480, max_lines=1024, max_batch=2, num_slots=1) as c:
        with pytest.raises(vpl.capi.VplError):
            c.run_resident(0, k=1)                     # nothing resident
        c.frontend_batch(frames[:2], scale=2, num_octaves=1, k=1)
        with pytest.raises(vpl.capi.VplError):
            c.run_resident(0, k=9)                     # k > max_k would overrun the match rows
        with pytest.raises(vpl.capi.VplError):
            c.run_resident(0, k=-1)
        c.run_resident(0, k=2)
        c.sync()
