"""GPU parity tests proper: every stage of the CUDA path (through the C ABI) against the
CPU oracle on the same inputs.  Bars: bit-exact for the integer/byte/index stages
(image primitives, ordering, Hamming), bit-exact LBD on identical keylines, and for LSD
the north-star bar (one-to-one, endpoints within 0.5 px for >= 99 %) -- in practice the
engine reproduces the sequential algorithm exactly, which the tests also record."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(vpl):
    c = vpl.Context(max_width=1280, max_height=720, max_octaves=2, max_lines=4096, max_batch=8, num_slots=2,
                    profile=True)
    yield c
    c.close()


def kl_fields_equal(a, b):
    return len(a) == len(b) and all(np.array_equal(a[n], b[n]) for n in a.dtype.names)


def match_segments(got, exp, tol=0.5):
    """one-to-one greedy match of segments by endpoint distance; fraction of exp matched within tol."""
    if len(exp) == 0:
        return 1.0 if len(got) == 0 else 0.0
    used = np.zeros(len(got), bool)
    ok = 0
    for e in exp:
        d = np.maximum(np.hypot(got[:, 0] - e[0], got[:, 1] - e[1]), np.hypot(got[:, 2] - e[2], got[:, 3] - e[3]))
        d[used] = np.inf
        j = int(np.argmin(d)) if len(d) else -1
        if j >= 0 and d[j] <= tol:
            used[j] = True
            ok += 1
    return ok / len(exp)


@pytest.mark.parametrize("shape", [(480, 752), (61, 83), (720, 1280), (100, 36)])
def test_image_primitives_bit_exact(ctx, orc, shape):
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(ctx.debug_stage(0, img), orc.gaussian_blur5(img))
    assert np.array_equal(ctx.debug_stage(1, img), orc.pyrdown(img))
    dx, dy = ctx.debug_stage(2, img)
    odx, ody = orc.sobel3(img)
    assert np.array_equal(dx, odx) and np.array_equal(dy, ody)
    assert np.array_equal(ctx.debug_stage(3, img), orc.resize08(orc.gaussian_blur7(img)))


def test_lsd_stages_bit_exact(ctx, orc, mh04):
    img = mh04[0]
    scaled, ang, order = orc.lsd_stages(img)
    assert np.array_equal(ctx.debug_stage(3, img), scaled)
    assert np.array_equal(ctx.debug_stage(4, img), ang)
    assert np.array_equal(ctx.debug_stage(5, img), order)


@pytest.mark.parametrize("k", [1, 5, 10])
def test_lsd_raw_vs_oracle_and_cv2_golden(ctx, orc, golden, mh04, k):
    blurred = orc.gaussian_blur5(mh04[k - 1])
    got = ctx.lsd_raw(blurred)
    seg, width, prec, nfa = orc.lsd_detect(blurred, refine=2)
    g = np.stack([got["x1"], got["y1"], got["x2"], got["y2"]], 1)
    # north-star bar
    assert len(g) == len(seg)
    assert match_segments(g, seg, 0.5) >= 0.99
    # what the engine actually delivers: the sequential result, bit for bit
    assert np.array_equal(g, seg)
    assert np.array_equal(g, golden["cv2_lsd"][f"adv{k}_lines"])
    assert np.allclose(got["width"], width, rtol=1e-12, atol=0)
    assert np.array_equal(got["prec"], prec)
    assert np.allclose(got["nfa"], nfa, rtol=1e-9, atol=1e-9)


def test_lsd_raw_synthetic_and_degenerate(ctx, orc, synth):
    for seed in (3, 4):
        img = synth.sequence(1, w=640, h=400, seed=seed, n_quads=25, n_strokes=40)[0]
        got = ctx.lsd_raw(img)
        seg = orc.lsd_detect(img, refine=2)[0]
        assert np.array_equal(np.stack([got["x1"], got["y1"], got["x2"], got["y2"]], 1), seg)
    assert len(ctx.lsd_raw(np.zeros((64, 64), np.uint8))) == 0
    assert len(ctx.lsd_raw(np.full((40, 50), 255, np.uint8))) == 0
    step = np.full((200, 200), 50, np.uint8); step[:, 100:] = 200
    got = ctx.lsd_raw(step)
    assert len(got) == 1
    assert np.array_equal(np.array([got["x1"][0], got["y1"][0], got["x2"][0], got["y2"][0]]),
                          orc.lsd_detect(step, refine=2)[0][0])


@pytest.mark.parametrize("octaves", [1, 2])
def test_lsd_detector_keylines_bit_exact(ctx, orc, mh04, octaves):
    frames = mh04[3:7]
    got = ctx.lsd_detect_batch(frames, scale=2, num_octaves=octaves)
    for f, img in enumerate(frames):
        exp = orc.lsd_detector_detect(img, scale=2, num_octaves=octaves)
        assert kl_fields_equal(got[f], exp), f"frame {f}"


def test_lbd_bit_exact_on_identical_keylines(ctx, orc, mh04):
    frames = mh04[0:3]
    kls = [orc.lsd_detector_detect(img, scale=2, num_octaves=2) for img in frames]
    got = ctx.lbd_compute_batch(frames, kls)
    for f, img in enumerate(frames):
        assert np.array_equal(got[f], orc.lbd_compute(img, kls[f])), f"frame {f}"


def test_lbd_float_descriptors_bit_exact(ctx, orc, mh04):
    """returnFloatDescr=true: the 72-float LBD vectors, bit for bit."""
    img = mh04[1]
    kl = orc.lsd_detector_detect(img, 2, 2)
    _, fd = orc.lbd_compute(img, kl, return_float=True)
    got = ctx.lbd_compute_float_batch(img[None], [kl])[0]
    assert got.shape == fd.shape and got.tobytes() == fd.tobytes()


def test_lbd_edge_keylines(ctx, orc, mh04):
    """lines hugging the border / very short / long: the clamped sampling must agree."""
    img = mh04[2]
    kl = orc.lsd_detector_detect(img, scale=2, num_octaves=1)
    sel = np.argsort(kl["lineLength"])
    pick = np.concatenate([sel[:20], sel[-20:], np.argsort(kl["pt_y"])[:20], np.argsort(-kl["pt_x"])[:20]])
    k = kl[np.unique(pick)]
    assert np.array_equal(ctx.lbd_compute_batch(img[None], [k])[0], orc.lbd_compute(img, k))


@pytest.mark.parametrize("k", [1, 2, 3])
def test_hamming_bit_exact(ctx, orc, golden, k):
    g = golden["cv2_hamming"]
    rng = np.random.default_rng(9)
    qs = [g["q"], rng.integers(0, 4, (300, 32), dtype=np.uint8), rng.integers(0, 256, (1, 32), dtype=np.uint8),
          rng.integers(0, 256, (33, 32), dtype=np.uint8)]
    ts = [g["t"], rng.integers(0, 4, (513, 32), dtype=np.uint8), rng.integers(0, 256, (700, 32), dtype=np.uint8),
          rng.integers(0, 256, (2, 32), dtype=np.uint8)]
    got = ctx.match_batch(qs, ts, k=k)
    for q, t, m in zip(qs, ts, got):
        idx, dist = orc.hamming_knn(q, t, k)
        assert np.array_equal(m["trainIdx"], idx)
        valid = idx >= 0
        assert np.array_equal(m["distance"][valid].astype(np.int32), dist[valid])
        assert np.array_equal(m["queryIdx"], np.repeat(np.arange(len(q))[:, None], k, 1))
    if k == 3:
        assert np.array_equal(got[0]["trainIdx"], g["idx"])  # cv2.BFMatcher golden, incl. planted ties


def test_frontend_fused_equals_oracle_chain(ctx, orc, mh04):
    """C1: the bundled mh04 frames, detect + describe + match(t, t-1), batches chained."""
    frames = mh04[:10]
    fe = __import__("vplines_slam_b200").FrontEnd(ctx, scale=2, num_octaves=1, k=2)
    kls, descs, ms = fe.run(frames)           # max_batch=8 -> two batches, chained
    prev = None
    for f, img in enumerate(frames):
        ekl = orc.lsd_detector_detect(img, 2, 1)
        edesc = orc.lbd_compute(img, ekl)
        assert kl_fields_equal(kls[f], ekl), f"keylines frame {f}"
        assert np.array_equal(descs[f], edesc), f"descriptors frame {f}"
        if prev is None:
            assert (ms[f]["trainIdx"] == -1).all()
        else:
            idx, dist = orc.hamming_knn(edesc, prev, 2)
            assert np.array_equal(ms[f]["trainIdx"], idx), f"matches frame {f}"
            assert np.array_equal(ms[f]["distance"].astype(np.int32), dist)
        prev = edesc


def test_frontend_grouped_driver_equals_plain_driver(vpl, mh04):
    """FrontEnd.run_grouped (frames pinned once, uploaded ahead, group submits, dense collects) on the device: what
    run() delivers (itself checked against the oracle chain above), whole and as two shards with the one-frame halo."""
    frames = np.ascontiguousarray(mh04[:11])
    with vpl.Context(max_width=752, max_height=480, max_octaves=1, max_lines=4096, max_batch=3, num_slots=2) as c:
        fe = vpl.FrontEnd(c, scale=2, num_octaves=1, k=2)
        ref = fe.run(frames)
        got = fe.run_grouped(frames)
        part = ([], [], [])
        for r in range(2):
            s, e, halo = vpl.shard_range(len(frames), r, 2)
            out = fe.run_grouped(frames, s, e, halo)
            for a, b in zip(part, out):
                a.extend(b)
    for a in (got, part):
        for x, y in zip(ref, a):
            assert len(x) == len(y) == len(frames)
            assert all(p.tobytes() == q.tobytes() for p, q in zip(x, y))


def test_screenshot_pair_5_10(ctx, orc, mh04):
    """C1's (5,10) pair: same pipeline through the OpenCV-shaped surface."""
    import vplines_slam_b200 as v
    det, bd, bm = v.LSDDetector.createLSDDetector(), v.BinaryDescriptor.createBinaryDescriptor(), \
        v.BinaryDescriptorMatcher.createBinaryDescriptorMatcher()
    k5 = det.detect(mh04[4], 2, 1, as_records=True); k10 = det.detect(mh04[9], 2, 1, as_records=True)
    d5 = bd.compute(mh04[4], k5); d10 = bd.compute(mh04[9], k10)
    m = bm.match(d10, d5)
    idx, dist = orc.hamming_knn(orc.lbd_compute(mh04[9], orc.lsd_detector_detect(mh04[9], 2, 1)),
                                orc.lbd_compute(mh04[4], orc.lsd_detector_detect(mh04[4], 2, 1)), 1)
    assert [x.trainIdx for x in m] == list(idx[:, 0]) and [int(x.distance) for x in m] == list(dist[:, 0])


def test_batch_invariance_and_determinism(ctx, synth):
    """Results do not depend on batching (size-independent property used at full size)."""
    frames = synth.config_sequence("C2_euroc_752x480", 6)
    a = ctx.lsd_detect_batch(frames)
    b = [ctx.lsd_detect_batch(frames[i:i + 1])[0] for i in range(len(frames))]
    c = ctx.lsd_detect_batch(frames)
    for x, y, z in zip(a, b, c):
        assert kl_fields_equal(x, y) and kl_fields_equal(x, z)


def test_two_octave_synthetic_c3(ctx, orc, synth):
    frames = synth.config_sequence("C3_d455_1280x720", 2)
    kls, descs, ms = ctx.frontend_batch(frames, scale=2, num_octaves=2, k=2)
    for f, img in enumerate(frames):
        ekl = orc.lsd_detector_detect(img, 2, 2)
        assert kl_fields_equal(kls[f], ekl)
        assert np.array_equal(descs[f], orc.lbd_compute(img, ekl))
    idx, dist = orc.hamming_knn(descs[1], descs[0], 2)
    assert np.array_equal(ms[1]["trainIdx"], idx)


def test_registered_host_buffer_path(ctx, vpl, synth):
    """vpl_host_register: frames uploaded straight from pinned caller memory give the same
    results as the staged path; d2h is the dense size."""
    frames = np.ascontiguousarray(synth.config_sequence("C2_euroc_752x480", 5))
    a = ctx.frontend_batch(frames, k=2)
    ctx.host_register(frames)
    try:
        b = ctx.frontend_batch(frames, k=2)
    finally:
        ctx.host_unregister(frames)
    for x, y in zip(a, b):
        for p, q in zip(x, y):
            assert p.tobytes() == q.tobytes()
    total = sum(len(k) for k in a[0])
    assert total * (68 + 32 + 2 * 16) <= ctx.last_d2h_bytes(0) + ctx.last_d2h_bytes(1)


def test_c4_manhattan_1080p_full_size(vpl, orc, synth):
    """C4 at full size: 1920x1080 Manhattan scene (~2000 lines).  One frame against the
    oracle bit for bit; the 2000x2000 brute-force match through size-independent properties
    (self-match: distance 0 at the lowest duplicate index; k=2 ordering) and against the oracle."""
    frames = synth.config_sequence("C4_manhattan_1920x1080", 2)
    with vpl.Context(max_width=1920, max_height=1080, max_octaves=1, max_lines=6144, max_batch=2, num_slots=2) as c:
        kls, descs, ms = c.frontend_batch(frames, scale=2, num_octaves=1, k=2)
        ekl = orc.lsd_detector_detect(frames[0], 2, 1)
        assert 1200 < len(ekl) < 6144
        assert kl_fields_equal(kls[0], ekl)
        assert np.array_equal(descs[0], orc.lbd_compute(frames[0], ekl))
        idx, dist = orc.hamming_knn(descs[1], descs[0], 2)
        assert np.array_equal(ms[1]["trainIdx"], idx)
        assert np.array_equal(ms[1]["distance"].astype(np.int32), dist)
        # self match: every code finds itself (or an identical earlier code) at distance 0
        m = c.match_batch([descs[0]], [descs[0]], k=2)[0]
        assert (m["distance"][:, 0] == 0).all()
        assert (m["trainIdx"][:, 0] <= np.arange(len(descs[0]))).all()
        assert (m["distance"][:, 0] <= m["distance"][:, 1]).all()
        # determinism / batch invariance at full size
        again = c.lsd_detect_batch(frames[1:2])[0]
        assert kl_fields_equal(again, kls[1])


def test_dense_collect_equals_frame_major(ctx, vpl, synth):
    frames = np.ascontiguousarray(synth.config_sequence("C2_euroc_752x480", 4))
    kls, descs, ms = ctx.frontend_batch(frames, k=1)
    n = ctx.submit(0, frames, k=1, chain=False)
    cap = ctx.max_lines
    counts = np.zeros(n, np.int32)
    kl = np.zeros(n * cap, vpl.capi.KEYLINE_DTYPE); d = np.zeros((n * cap, 32), np.uint8)
    m = np.zeros((n * cap, 1), vpl.capi.DMATCH_DTYPE)
    ctx.host_register(d)
    try:
        total = ctx.collect_dense_into(0, counts, kl, d, m)
    finally:
        ctx.host_unregister(d)
    assert total == sum(len(k) for k in kls) and list(counts) == [len(k) for k in kls]
    off = 0
    for f in range(n):
        c = counts[f]
        assert kl[off:off + c].tobytes() == kls[f].tobytes()
        assert np.array_equal(d[off:off + c], descs[f])
        assert m[off:off + c].tobytes() == ms[f].tobytes()
        off += c


def test_three_octaves_and_unblurred_variant(vpl, orc, mh04):
    img = mh04[7]
    with vpl.Context(max_width=752, max_height=480, max_octaves=3, max_lines=4096, max_batch=2) as c:
        got = c.lsd_detect_batch(img[None], scale=2, num_octaves=3)[0]
        exp = orc.lsd_detector_detect(img, 2, 3)
        assert set(np.unique(exp["octave"])) == {0, 1, 2}
        assert kl_fields_equal(got, exp)
        assert np.array_equal(c.lbd_compute_batch(img[None], [exp])[0], orc.lbd_compute(img, exp))
    # LSDDetector variant with the pyramid blur commented out (VplConfig.blur_first = 0); BinaryDescriptor
    # still blurs its own pyramid
    with vpl.Context(max_width=752, max_height=480, max_octaves=2, max_lines=4096, max_batch=2, blur_first=False) as c:
        kls, descs, _ = c.frontend_batch(img[None], scale=2, num_octaves=2, k=1)
        exp = orc.lsd_detector_detect(img, 2, 2, blur_first=False)
        assert kl_fields_equal(kls[0], exp)
        assert np.array_equal(descs[0], orc.lbd_compute(img, exp))


def test_strided_input_and_knn_k5(ctx, vpl, orc, mh04):
    import ctypes as C
    big = np.zeros((480, 800), np.uint8)
    big[:, :752] = mh04[2]
    view = big[:, :752]                      # row pitch 800 bytes
    L = vpl.capi.load()
    cap = 4096
    kl = np.zeros(cap, vpl.capi.KEYLINE_DTYPE); cnt = np.zeros(1, np.int32)
    ptrs = (C.c_void_p * 1)(view.ctypes.data)
    r = L.vpl_lsd_detect_batch(ctx._h, ptrs, 1, 752, 480, 800, 2, 1, kl.ctypes.data_as(C.c_void_p),
                               cnt.ctypes.data_as(C.c_void_p), cap)
    assert r == 0
    exp = orc.lsd_detector_detect(mh04[2], 2, 1)
    assert kl_fields_equal(kl[:cnt[0]], exp)
    d = orc.lbd_compute(mh04[2], exp)
    m = ctx.match_batch([d[:300]], [d[100:500]], k=5)[0]
    idx, dist = orc.hamming_knn(d[:300], d[100:500], 5)
    assert np.array_equal(m["trainIdx"], idx) and np.array_equal(m["distance"].astype(np.int32), dist)


def test_capacity_errors_are_loud(vpl, mh04):
    with vpl.Context(max_width=752, max_height=480, max_octaves=1, max_lines=64, max_batch=2) as c:
        with pytest.raises(vpl.VplError) as e:
            c.lsd_detect_batch(mh04[:1])     # ~800 lines > max_lines
        assert e.value.code == vpl.capi.VPL_E_CAPACITY
        # the context stays usable afterwards
        assert len(c.lsd_detect_batch(np.zeros((1, 480, 752), np.uint8))[0]) == 0


def test_preprocess_remap_clahe_bit_exact(vpl, orc, mh04):
    """undistortion remap + CLAHE(3.0, 8x8) on the device, and the front end fed by it,
    against the oracle (which is pinned to cv2.remap / cv2.CLAHE)."""
    from test_oracle_preproc import euroc_maps
    mapx, mapy = euroc_maps()
    frames = mh04[:3]
    with vpl.Context(max_width=752, max_height=480, max_octaves=1, max_lines=4096, max_batch=4) as c:
        c.set_preprocess(mapx, mapy, clahe_clip=3.0, clahe_tiles=8)
        got = c.preprocess_batch(frames)
        exp = [orc.clahe(orc.remap_linear(f, mapx, mapy), 3.0, 8) for f in frames]
        for g, e in zip(got, exp):
            assert np.array_equal(g, e)
        kls, descs, _ = c.frontend_batch(frames, k=1)
        for f in range(3):
            ekl = orc.lsd_detector_detect(exp[f], 2, 1)
            assert kl_fields_equal(kls[f], ekl)
            assert np.array_equal(descs[f], orc.lbd_compute(exp[f], ekl))
        # remap only / CLAHE only / off again
        c.set_preprocess(mapx, mapy, clahe_clip=0.0)
        assert np.array_equal(c.preprocess_batch(frames[:1])[0], orc.remap_linear(frames[0], mapx, mapy))
        c.set_preprocess(None, None, size=(752, 480), clahe_clip=3.0, clahe_tiles=8)
        assert np.array_equal(c.preprocess_batch(frames[:1])[0], orc.clahe(frames[0], 3.0, 8))
        c.set_preprocess(None, None, size=(752, 480), clahe_clip=0.0)
        assert np.array_equal(c.preprocess_batch(frames[:1])[0], frames[0])
    # sizes that are not a multiple of the CLAHE grid, jittered maps with exact-integer coordinates
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, (75, 100), dtype=np.uint8)
    mx = (np.arange(100)[None, :] + rng.uniform(-4, 4, (75, 100))).astype(np.float32)
    my = (np.arange(75)[:, None] + rng.uniform(-4, 4, (75, 100))).astype(np.float32)
    mx[::5, ::3] = np.round(mx[::5, ::3]); my[::5, ::3] = np.round(my[::5, ::3])
    with vpl.Context(max_width=100, max_height=75, max_octaves=1, max_lines=512, max_batch=2) as c:
        c.set_preprocess(mx, my, clahe_clip=2.0, clahe_tiles=4)
        assert np.array_equal(c.preprocess_batch(img[None])[0], orc.clahe(orc.remap_linear(img, mx, my), 2.0, 4))


def test_errors(ctx, vpl):
    with pytest.raises(vpl.VplError):
        ctx.lsd_detect_batch(np.zeros((1, 2000, 2000), np.uint8))       # larger than the context
    with pytest.raises(vpl.VplError):
        ctx.lsd_detect_batch(np.zeros((64, 64, 64), np.uint8))          # batch > max_batch
    with pytest.raises(RuntimeError):
        vpl.LSDDetector.createLSDDetector().detect(np.zeros((64, 64), np.float32), 2, 1)
    assert vpl.BinaryDescriptorMatcher().match(np.zeros((0, 32), np.uint8), np.zeros((4, 32), np.uint8)) == []


def _bench_path(vpl, frames, n_batches, k, octaves, cap, upload_ahead=False, group=False, profile=True):
    """The exact call sequence bench.py times end to end: caller frames pinned with vpl_host_register, batches
    submitted to alternating slots with chaining (the frames of batch i + 2 uploaded ahead with vpl_frontend_upload
    while batch i runs, if asked), results collected in dense form into registered buffers."""
    frames = np.ascontiguousarray(frames)
    n, h, w = frames.shape
    per = (n + n_batches - 1) // n_batches
    S = 2
    outs = []
    with vpl.Context(max_width=w, max_height=h, max_octaves=octaves, max_lines=cap, max_batch=per, num_slots=S,
                     blur_first=True, profile=profile) as c:
        c.host_register(frames)
        rows = per * cap
        kl = [np.zeros(rows, vpl.capi.KEYLINE_DTYPE) for _ in range(S)]
        counts = [np.zeros(per, np.int32) for _ in range(S)]
        desc = [np.zeros((rows, 32), np.uint8) for _ in range(S)]
        mt = [np.zeros((rows, k), vpl.capi.DMATCH_DTYPE) for _ in range(S)]
        for a in kl + desc + mt:
            c.host_register(a)
        pending = []

        def collect():
            ps, lo, hi = pending.pop(0)
            total = c.collect_dense_into(ps, counts[ps], kl[ps], desc[ps], mt[ps])
            outs.append((lo, hi, counts[ps][:hi - lo].copy(), kl[ps][:total].copy(), desc[ps][:total].copy(),
                         mt[ps][:total].copy()))

        def rng(i):
            return i * per, min(n, (i + 1) * per)

        if upload_ahead:
            for i in range(min(S, n_batches)):
                c.upload(i % S, frames[slice(*rng(i))])
        if group:
            # bench.py's default: the slots' uploaded batches submitted as one group; group g is submitted BEFORE group
            # g-1 is collected (two result generations per slot), or after it (group == "collect_first")
            groups = [list(range(g, min(g + S, n_batches))) for g in range(0, n_batches, S)]

            def submit(grp):
                c.submit_group([i % S for i in grp], [rng(i)[1] - rng(i)[0] for i in grp], w, h, scale=2,
                               num_octaves=octaves, k=k, chain=[i > 0 for i in grp])
                pending.extend((i % S,) + rng(i) for i in grp)

            submit(groups[0])
            for gi in range(1, len(groups)):
                for i in groups[gi]:
                    c.upload(i % S, frames[slice(*rng(i))])
                if group == "collect_first":
                    for _ in groups[gi - 1]:
                        collect()
                    submit(groups[gi])
                else:
                    submit(groups[gi])
                    for _ in groups[gi - 1]:
                        collect()
            n_batches = 0
        for i in range(n_batches):
            s = i % S
            if len(pending) == S:
                collect()
            lo, hi = rng(i)
            if upload_ahead:
                c.submit_uploaded(s, hi - lo, w, h, scale=2, num_octaves=octaves, k=k, chain=(i > 0))
                if i + S < n_batches:
                    c.upload(s, frames[slice(*rng(i + S))])
            else:
                c.submit(s, frames[lo:hi], scale=2, num_octaves=octaves, k=k, chain=(i > 0))
            pending.append((s, lo, hi))
        while pending:
            collect()
    return outs


@pytest.mark.parametrize("name,n,batches,k,octaves,cap,ahead", [("C2_euroc_752x480", 33, 3, 1, 1, 1024, "no"),
                                                                ("C2_euroc_752x480", 33, 5, 1, 1, 1024, "ahead"),
                                                                ("C2_euroc_752x480", 33, 5, 1, 1, 1024, "group"),
                                                                ("C2_euroc_752x480", 33, 9, 1, 1, 1024, "group"),
                                                                ("C2_euroc_752x480", 33, 5, 1, 1, 1024, "collect_first"),
                                                                ("C3_d455_1280x720", 9, 3, 2, 2, 2048, "group"),
                                                                ("C3_d455_1280x720", 9, 3, 2, 2, 2048, "no")])
def test_bench_path_against_oracle(vpl, orc, synth, name, n, batches, k, octaves, cap, ahead):
    """The headline configs through bench.py's own path (host_register + submit(chain) over three batches +
    collect_dense; with upload-ahead; with upload-ahead and group submits = the bench's default) against the oracle chain: every KeyLine, descriptor and match of every frame bit-equal; the
    first frame of a later batch is matched against the last frame of the batch before it."""
    import bench
    frames = np.ascontiguousarray(synth.config_sequence(name, n))
    # (stage events off for the queued groups, as in bench.py's timed pass: with them on a submit waits for the slot)
    outs = _bench_path(vpl, frames, batches, k, octaves, cap, ahead != "no",
                       ahead if ahead in ("group", "collect_first") else False, profile=(ahead != "group"))
    assert [o[0] for o in outs] == [i * ((n + batches - 1) // batches) for i in range(batches)]
    n_lines = 0
    for lo, hi, counts, kl, desc, mt in outs:
        prev = frames[lo - 1] if lo > 0 else None
        bad = bench.verify_dense_step(frames[lo:hi], prev, counts, kl, desc, mt, k, octaves, list(range(hi - lo)))
        assert bad == [], (name, lo, bad)
        n_lines += int(counts.sum())
    assert n_lines > 50 * n


def test_upload_ahead_argument_errors(vpl, synth):
    frames = np.ascontiguousarray(synth.config_sequence("C2_euroc_752x480", 4))
    with vpl.Context(max_width=752, max_height=480, max_octaves=1, max_lines=1024, max_batch=4, num_slots=2) as c:
        with pytest.raises(vpl.capi.VplError, match="no uploaded batch"):
            c.submit_uploaded(0, 4, 752, 480)
        with pytest.raises(vpl.capi.VplError, match="vpl_host_register"):
            c.upload(0, frames)  # not pinned
        c.host_register(frames)
        c.upload(0, frames[:3])
        with pytest.raises(vpl.capi.VplError, match="uploaded batch of 3 frames"):
            c.submit_uploaded(0, 4, 752, 480)
        # a group is submitted whole or not at all: slot 1 has nothing uploaded, slot 0 keeps its batch
        with pytest.raises(vpl.capi.VplError, match="slot 1 holds no uploaded batch"):
            c.submit_group([0, 1], [3, 3], 752, 480)
        with pytest.raises(vpl.capi.VplError, match="twice in the group"):
            c.submit_group([0, 0], [3, 3], 752, 480)
        with pytest.raises(vpl.capi.VplError, match="already holds an uploaded batch"):
            c.upload(0, frames[:2])
        c.submit_uploaded(0, 3, 752, 480, k=0)
        kl = np.zeros(3 * 1024, vpl.capi.KEYLINE_DTYPE); counts = np.zeros(3, np.int32)
        d = np.zeros((3 * 1024, 32), np.uint8); m = np.zeros((3 * 1024, 1), vpl.capi.DMATCH_DTYPE)
        total = c.collect_dense_into(0, counts, kl, d, m)
        ref = c.frontend_batch(frames[:3], k=0)[0]
        assert total == sum(len(r) for r in ref) and kl[:len(ref[0])].tobytes() == ref[0].tobytes()
        # two batches queued on one slot: the second one only from an upload, no third, collects come oldest first
        c.upload(0, frames[:2]); c.submit_uploaded(0, 2, 752, 480, k=0)
        with pytest.raises(vpl.capi.VplError, match="still in flight"):
            c.submit(0, frames[:1], k=0)
        c.upload(0, frames[2:4]); c.submit_uploaded(0, 2, 752, 480, k=0)
        with pytest.raises(vpl.capi.VplError, match="two uncollected batches"):
            c.upload(0, frames[:1])
        t1 = c.collect_dense_into(0, counts, kl, d, m); first = kl[:t1].copy()
        t2 = c.collect_dense_into(0, counts, kl, d, m); second = kl[:t2].copy()
        with pytest.raises(vpl.capi.VplError, match="no front-end batch in flight"):
            c.collect_dense_into(0, counts, kl, d, m)
        ref4 = c.frontend_batch(frames, k=0)[0]
        assert first.tobytes() == np.concatenate(ref4[:2]).tobytes() and second.tobytes() == np.concatenate(ref4[2:]).tobytes()


@pytest.mark.parametrize("cap", [64, 512, 4096])
def test_engine_ring_capacity_fallbacks(vpl, orc, mh04, synth, cap):
    """The speculative region engine with rings so small that its fallbacks run all the time (a region that does
    not fit a lane's ring undoes what the lane has parked; one that does not fit at all runs alone with the whole
    arena): KeyLines still equal the oracle's bit for bit."""
    frames = np.stack([mh04[0], synth.config_sequence("C2_euroc_752x480", 1)[0]])
    with vpl.Context(max_width=752, max_height=480, max_octaves=1, max_lines=4096, max_batch=2, num_slots=1) as c:
        c.set_engine(1)
        c.set_engine_ring_cap(cap)
        kls = c.lsd_detect_batch(frames)
        c.set_engine_ring_cap(0)
        again = c.lsd_detect_batch(frames)
    for f in range(2):
        ekl = orc.lsd_detector_detect(frames[f], 2, 1)
        assert kl_fields_equal(kls[f], ekl), (cap, f, len(kls[f]), len(ekl))
        assert kl_fields_equal(again[f], ekl), f


@pytest.mark.parametrize("kind", [0, 1])
def test_both_region_engines_against_oracle(vpl, orc, mh04, synth, kind):
    """The default (warp-cooperative, sequential seed order) and the speculative (32 seeds in flight) region engine both
    reproduce the oracle's KeyLines bit for bit: real EuRoC frames, synthetic C2 frames, two octaves of a C3 frame."""
    frames = np.concatenate([mh04[:6], synth.config_sequence("C2_euroc_752x480", 4)])
    with vpl.Context(max_width=1280, max_height=720, max_octaves=2, max_lines=4096, max_batch=40, num_slots=1) as c:
        c.set_engine(kind)
        kls = c.lsd_detect_batch(frames)
        for f in range(len(frames)):
            assert kl_fields_equal(kls[f], orc.lsd_detector_detect(frames[f], 2, 1)), (kind, f)
        c3 = synth.config_sequence("C3_d455_1280x720", 1)
        k3 = c.lsd_detect_batch(c3, scale=2, num_octaves=2)
        assert kl_fields_equal(k3[0], orc.lsd_detector_detect(c3[0], 2, 2)), kind
        many = np.ascontiguousarray(np.concatenate([frames] * 4))
        km = c.lsd_detect_batch(many)
        for f in range(len(many)):
            assert kl_fields_equal(km[f], kls[f % len(frames)]), (kind, f)


@pytest.mark.parametrize("variant", [0, 2, 3])
def test_region_engine_builds_selected_by_launch_size(variant):
    """The register budgets of the default engine (1 / 2 / 4 frames per block; the 32-register build is what large
    launches take) are the same algorithm: forced through VPL_ENGINE_VARIANT in a process of their own, each gives the
    oracle's KeyLines bit for bit."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, importlib, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import vplines_slam_b200 as vpl\n"
        "from oracle import oracle as O\n"
        "O.build(); O.lib()\n"
        "synth = importlib.import_module('vplines-slam_b200.synth')\n"
        "frames = np.ascontiguousarray(synth.config_sequence('C2_euroc_752x480', 5))\n"
        "with vpl.Context(max_width=752, max_height=480, max_octaves=1, max_lines=4096, max_batch=8, num_slots=1) as c:\n"
        "    kls = c.lsd_detect_batch(frames)\n"
        "for f in range(len(frames)):\n"
        "    e = O.lsd_detector_detect(frames[f], 2, 1)\n"
        "    assert len(e) > 50 and kls[f].tobytes() == e.tobytes(), f\n"
        "print('ok')\n" % root)
    env = dict(os.environ, VPL_ENGINE_VARIANT=str(variant))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd=root)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
