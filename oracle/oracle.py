"""ctypes loader for the CPU oracle (oracle/_build/libvpl_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvpl_oracle.so")

KEYLINE_DTYPE = np.dtype(
    [("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"), ("pt_y", "<f4"),
     ("response", "<f4"), ("size", "<f4"), ("startPointX", "<f4"), ("startPointY", "<f4"),
     ("endPointX", "<f4"), ("endPointY", "<f4"), ("sPointInOctaveX", "<f4"),
     ("sPointInOctaveY", "<f4"), ("ePointInOctaveX", "<f4"), ("ePointInOctaveY", "<f4"),
     ("lineLength", "<f4"), ("numOfPixels", "<i4")])
assert KEYLINE_DTYPE.itemsize == 68


def _cpu_stamp():
    """ISA flags of this host: the library is built -march=native, so it must be rebuilt on another CPU."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((l.split(":", 1)[1] for l in f if l.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha1(" ".join(sorted(flags.split())).encode()).hexdigest()


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    stamp = os.path.join(_HERE, "_build", "cpu.stamp")
    same_cpu = os.path.exists(stamp) and open(stamp).read().strip() == _cpu_stamp()
    if (not force and same_cpu and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    with open(stamp, "w") as f:
        f.write(_cpu_stamp())
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()  # no-op when the library is current for this CPU
        L = ctypes.CDLL(_SO)
        L.orc_fast_atan2.restype = ctypes.c_float
        L.orc_fast_atan2.argtypes = [ctypes.c_float, ctypes.c_float]
        L.orc_lsd_detect.restype = ctypes.c_int
        L.orc_lsd_detector_detect.restype = ctypes.c_int
        L.orc_lsd_stages.restype = ctypes.c_int
        L.orc_lsd_candidates.restype = ctypes.c_int
        L.orc_lbd_compute.restype = ctypes.c_int
        L.orc_frontend_sequence.restype = ctypes.c_int64
        L.orc_frontend_sequence_mt.restype = ctypes.c_int64
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _u8(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img


def gaussian_blur5(img):
    img = _u8(img); out = np.empty_like(img)
    lib().orc_gaussian_blur5(_p(img), img.shape[1], img.shape[0], _p(out))
    return out


def gaussian_blur7(img):
    img = _u8(img); out = np.empty_like(img)
    lib().orc_gaussian_blur7_s075(_p(img), img.shape[1], img.shape[0], _p(out))
    return out


def resize08(img):
    img = _u8(img); h, w = img.shape
    dw, dh = ctypes.c_int(), ctypes.c_int()
    lib().orc_resize_08(_p(img), w, h, None, ctypes.byref(dw), ctypes.byref(dh))
    out = np.empty((dh.value, dw.value), np.uint8)
    lib().orc_resize_08(_p(img), w, h, _p(out), ctypes.byref(dw), ctypes.byref(dh))
    return out


def pyrdown(img):
    img = _u8(img); h, w = img.shape
    out = np.empty((h // 2, w // 2), np.uint8)
    lib().orc_pyrdown_half(_p(img), w, h, _p(out))
    return out


def sobel3(img):
    img = _u8(img); h, w = img.shape
    dx = np.empty((h, w), np.int16); dy = np.empty((h, w), np.int16)
    lib().orc_sobel3(_p(img), w, h, _p(dx), _p(dy))
    return dx, dy


def remap_linear(img, mapx, mapy):
    """cv::remap(img, mapx, mapy, INTER_LINEAR) with float32 maps, constant-0 border."""
    img = _u8(img); h, w = img.shape
    mapx = np.ascontiguousarray(mapx, np.float32); mapy = np.ascontiguousarray(mapy, np.float32)
    dh, dw = mapx.shape
    out = np.empty((dh, dw), np.uint8)
    lib().orc_remap_linear(_p(img), w, h, _p(mapx), _p(mapy), dw, dh, _p(out))
    return out


def clahe(img, clip_limit=3.0, tiles=8):
    """cv::createCLAHE(clip_limit, (tiles, tiles)).apply(img)."""
    img = _u8(img); h, w = img.shape
    out = np.empty_like(img)
    lib().orc_clahe(_p(img), w, h, ctypes.c_double(clip_limit), int(tiles), _p(out))
    return out


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def lsd_detect(img, refine=2, scale08=True, cap=1 << 16):
    """cv::LineSegmentDetector::detect -> (seg[n,4] f32, width[n], prec[n], nfa[n])."""
    img = _u8(img); h, w = img.shape
    seg = np.zeros((cap, 4), np.float32)
    wd = np.zeros(cap); pr = np.zeros(cap); nf = np.zeros(cap)
    n = lib().orc_lsd_detect(_p(img), w, h, int(refine), int(bool(scale08)), _p(seg), _p(wd), _p(pr),
                             _p(nf), cap)
    n = min(n, cap)
    return seg[:n].copy(), wd[:n].copy(), pr[:n].copy(), nf[:n].copy()


def lsd_candidates(img, cap=1 << 16):
    img = _u8(img); h, w = img.shape
    out = np.zeros((cap, 16), np.float64)
    n = lib().orc_lsd_candidates(_p(img), w, h, _p(out), cap)
    return out[:min(n, cap)].copy()


def lsd_stages(img):
    """-> (scaled u8 [hs,ws], angle degrees f32 [hs,ws] (-1024 undefined), ordered defined pixel indices)."""
    img = _u8(img); h, w = img.shape
    ws, hs = ctypes.c_int(), ctypes.c_int()
    lib().orc_resize_08(_p(img), w, h, None, ctypes.byref(ws), ctypes.byref(hs))
    scaled = np.zeros((hs.value, ws.value), np.uint8)
    ang = np.zeros((hs.value, ws.value), np.float32)
    order = np.zeros(hs.value * ws.value, np.int32)
    n = lib().orc_lsd_stages(_p(img), w, h, _p(scaled), _p(ang), _p(order), ctypes.byref(ws), ctypes.byref(hs))
    return scaled, ang, order[:n].copy()


def lsd_detector_detect(img, scale=2, num_octaves=1, blur_first=True, cap=1 << 16):
    """LSDDetector::detect -> structured array of KeyLine."""
    img = _u8(img); h, w = img.shape
    kl = np.zeros(cap, KEYLINE_DTYPE)
    n = lib().orc_lsd_detector_detect(_p(img), w, h, int(scale), int(num_octaves), int(bool(blur_first)),
                                      _p(kl), cap)
    return kl[:min(n, cap)].copy()


def lbd_compute(img, keylines, return_float=False):
    """BinaryDescriptor::compute -> desc[n,32] u8 (and [n,72] f32 if return_float)."""
    img = _u8(img); h, w = img.shape
    kl = np.ascontiguousarray(keylines, dtype=KEYLINE_DTYPE)
    n = len(kl)
    desc = np.zeros((n, 32), np.uint8)
    fd = np.zeros((n, 72), np.float32)
    if n:
        lib().orc_lbd_compute(_p(img), w, h, _p(kl), n, _p(desc), _p(fd))
    return (desc, fd) if return_float else desc


def hamming_knn(q, t, k=1):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.full((len(q), k), -1, np.int32); dist = np.full((len(q), k), -1, np.int32)
    if len(q):
        lib().orc_hamming_knn(_p(q), len(q), _p(t), len(t), int(k), _p(idx), _p(dist))
    return idx, dist


# ---- the reference's real detector: EDLines (orc_edlines.c) -------------------------------
LINE_DTYPE = np.dtype([("endpoint", "<f4", 4), ("equation", "<f8", 3), ("center", "<f4", 2),
                       ("length", "<f4"), ("pad_", "<f4")])
assert LINE_DTYPE.itemsize == 56


class EDLineParam(ctypes.Structure):
    """EDLineParam (line_matching/src/edline_detector.h:32-40).  Defaults = the tracker node's
    (feature_tracker/src/line_feature_tracker_node.cpp:203 with the EuRoC yaml)."""
    _fields_ = [("ksize", ctypes.c_int), ("sigma", ctypes.c_float), ("gradientThreshold", ctypes.c_float),
                ("anchorThreshold", ctypes.c_float), ("scanIntervals", ctypes.c_int),
                ("minLineLen", ctypes.c_int), ("lineFitErrThreshold", ctypes.c_double)]

    def __init__(self, ksize=5, sigma=1.0, gradientThreshold=30, anchorThreshold=5, scanIntervals=2,
                 minLineLen=35, lineFitErrThreshold=1.8):
        super().__init__(ksize, sigma, gradientThreshold, anchorThreshold, scanIntervals, minLineLen,
                         lineFitErrThreshold)


def _edline_call(fn, img, param, smoothed, cap):
    img = _u8(img); h, w = img.shape
    out = np.zeros(cap, LINE_DTYPE)
    xy = np.zeros(2 * (w * h // 5) + 16, np.uint32); sid = np.zeros(w * h // 100 + 2, np.uint32)
    npx, nch = ctypes.c_int(), ctypes.c_int()
    return img, w, h, out, xy, sid, npx, nch


def edline_detect(img, param=None, smoothed=True, cap=1 << 14, stages=False):
    """EDLineDetector::EDline -> Line records (LINE_DTYPE) in (chain, position) order;
    with stages=True also (chain_xy, chain_sid)."""
    param = param or EDLineParam()
    img, w, h, out, xy, sid, npx, nch = _edline_call(None, img, param, smoothed, cap)
    L = lib(); L.orc_edline_detect.restype = ctypes.c_int
    n = L.orc_edline_detect(_p(img), w, h, ctypes.byref(param), int(bool(smoothed)), _p(out), cap, _p(xy), _p(sid),
                            ctypes.byref(npx), ctypes.byref(nch))
    lines = out[:min(n, cap)].copy()
    if stages:
        return lines, xy[:npx.value].copy(), sid[:nch.value + 1].copy()
    return lines


_REF_SO = os.path.join(_HERE, "_ref", "libref_linefront.so")
_ref = None


def ref_available():
    return os.path.exists(_REF_SO) or os.path.isdir("/root/reference/line_matching/src")


def build_ref():
    """oracle/_ref/libref_linefront.so = the reference's own edline_detector.cpp, line_matching.cpp and
    lk_tracker_invoker_2d.cpp compiled against oracle/cvshim.  Only possible where /root/reference exists; elsewhere the prebuilt file is used."""
    if os.path.isdir("/root/reference/line_matching/src"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return _REF_SO if os.path.exists(_REF_SO) else None


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_SO):
            build_ref()
        _ref = ctypes.CDLL(_REF_SO)
        _ref.ref_edline_detect.restype = ctypes.c_int
        _ref.ref_edline_sequence_mt.restype = ctypes.c_longlong
        _ref.ref_line_matching.restype = ctypes.c_int
        _ref.ref_linefront_sequence_mt.restype = ctypes.c_longlong
    return _ref


def ref_edline_detect(img, param=None, smoothed=True, cap=1 << 14, stages=False):
    """The same call through the reference's own code (oracle/_ref)."""
    param = param or EDLineParam()
    img, w, h, out, xy, sid, npx, nch = _edline_call(None, img, param, smoothed, cap)
    n = ref_lib().ref_edline_detect(_p(img), w, h, ctypes.byref(param), int(bool(smoothed)), _p(out), cap,
                                    _p(xy), len(xy), _p(sid), len(sid) - 1, ctypes.byref(npx), ctypes.byref(nch))
    lines = out[:min(n, cap)].copy()
    if stages:
        return lines, xy[:npx.value].copy(), sid[:nch.value + 1].copy()
    return lines


def edline_sequence(frames, param=None, smoothed=True, threads=1, use_ref=False):
    """Total lines over a frame stack, frames spread over host threads (CPU baseline timing)."""
    param = param or EDLineParam()
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    if use_ref:
        return int(ref_lib().ref_edline_sequence_mt(_p(frames), n, w, h, ctypes.byref(param),
                                                    int(bool(smoothed)), int(threads)))
    L = lib(); L.orc_edline_sequence_mt.restype = ctypes.c_int64
    return int(L.orc_edline_sequence_mt(_p(frames), n, w, h, ctypes.byref(param), int(bool(smoothed)),
                                        int(threads)))


# ---- the reference's real matcher: LineMatching::Matching (orc_linematch.c) ---------------------
class LineMatchParam(ctypes.Structure):
    _fields_ = [("step", ctypes.c_int), ("closest_line_threshold", ctypes.c_float),
                ("line_matching_ratio", ctypes.c_float), ("line_distance_error_ratio", ctypes.c_float),
                ("klt_error_threshold", ctypes.c_float), ("win", ctypes.c_int), ("max_level", ctypes.c_int),
                ("max_count", ctypes.c_int), ("epsilon", ctypes.c_double), ("min_eig", ctypes.c_float),
                ("topo_distance_threshold", ctypes.c_float), ("topo_length_ratio", ctypes.c_float),
                ("topo_violation_ratio", ctypes.c_float)]

    def __init__(self, **kw):
        super().__init__()
        lib().orc_lm_default_param(ctypes.byref(self))
        for k, v in kw.items():
            setattr(self, k, v)


def pyrdown_std(img):
    img = _u8(img); h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyrdown_std(_p(img), w, h, _p(out))
    return out


def klt_calc2d(img_ref, img_cur, pts, win=13, max_level=3, max_count=30, epsilon=0.001, min_eig=1e-4, illum=True):
    """KLT::calc2D (flags 0) -> (next_pts [n,2] f32, status u8, err f32)."""
    img_ref = _u8(img_ref); img_cur = _u8(img_cur); h, w = img_ref.shape
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    n = len(pts)
    nxt = np.zeros((n, 2), np.float32); st = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
    lib().orc_klt_calc2d(_p(img_ref), _p(img_cur), w, h, _p(pts), n, int(win), int(max_level), int(max_count),
                         ctypes.c_double(epsilon), ctypes.c_float(min_eig), int(bool(illum)), _p(nxt), _p(st), _p(err))
    return nxt, st, err


def line_matching(img_ref, img_cur, lines_ref, lines_cur, param=None, illum=True, topo=True, details=False):
    """LineMatching::Matching -> ref_to_cur int32[n_ref] (None when the reference returns false);
    details=True adds a dict with the per-anchor arrays."""
    param = param or LineMatchParam()
    img_ref = _u8(img_ref); img_cur = _u8(img_cur); h, w = img_ref.shape
    lr = np.ascontiguousarray(lines_ref, LINE_DTYPE); lc = np.ascontiguousarray(lines_cur, LINE_DTYPE)
    cap = int(sum(int(l / max(param.step, 1)) + 2 for l in lr["length"])) + 8 if len(lr) else 8
    r2c = np.full(len(lr), -1, np.int32)
    kr = np.zeros((cap, 2), np.float32); kc = np.zeros((cap, 2), np.float32)
    st = np.zeros(cap, np.uint8); er = np.zeros(cap, np.float32); k2l = np.zeros(cap, np.int32)
    nkp = ctypes.c_int()
    L = lib(); L.orc_line_matching.restype = ctypes.c_int
    ok = L.orc_line_matching(_p(img_ref), _p(img_cur), w, h, _p(lr), len(lr), _p(lc), len(lc), ctypes.byref(param),
                             int(bool(illum)), int(bool(topo)), _p(r2c), _p(kr), _p(kc), _p(st), _p(er), _p(k2l), cap,
                             ctypes.byref(nkp))
    res = r2c if ok else None
    if details:
        n = nkp.value
        return res, dict(kps_ref=kr[:n].copy(), kps_cur=kc[:n].copy(), status=st[:n].copy(), err=er[:n].copy(),
                         kp2line=k2l[:n].copy())
    return res


def ref_line_matching(img_ref, img_cur, lines_ref, lines_cur, illum=True, topo=True, details=False):
    """The same call through the reference's own line_matching.cpp / lk_tracker_invoker_2d.cpp (oracle/_ref)."""
    img_ref = _u8(img_ref); img_cur = _u8(img_cur); h, w = img_ref.shape
    lr = np.ascontiguousarray(lines_ref, LINE_DTYPE); lc = np.ascontiguousarray(lines_cur, LINE_DTYPE)
    cap = int(sum(int(l / 10) + 2 for l in lr["length"])) + 8 if len(lr) else 8
    r2c = np.full(len(lr), -1, np.int32)
    kr = np.zeros((cap, 2), np.float32); kc = np.zeros((cap, 2), np.float32)
    st = np.zeros(cap, np.uint8); er = np.zeros(cap, np.float32); k2l = np.zeros(cap, np.int32)
    nkp = ctypes.c_int()
    ok = ref_lib().ref_line_matching(_p(img_ref), _p(img_cur), w, h, _p(lr), len(lr), _p(lc), len(lc), int(bool(illum)),
                                     int(bool(topo)), _p(r2c), _p(kr), _p(kc), _p(st), _p(er), _p(k2l), cap,
                                     ctypes.byref(nkp))
    res = r2c if ok else None
    if details:
        n = nkp.value
        return res, dict(kps_ref=kr[:n].copy(), kps_cur=kc[:n].copy(), status=st[:n].copy(), err=er[:n].copy(),
                         kp2line=k2l[:n].copy())
    return res


def linefront_sequence(frames, param=None, smoothed=True, threads=1, use_ref=False):
    """EDLines on every frame + Matching(prev, cur) on every consecutive pair (the tracker's hot loop),
    frames over host threads in contiguous chunks with a one-frame halo -> number of matched lines."""
    param = param or EDLineParam()
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    if use_ref:
        return int(ref_lib().ref_linefront_sequence_mt(_p(frames), n, w, h, ctypes.byref(param), int(bool(smoothed)),
                                                       int(threads)))
    L = lib(); L.orc_linefront_sequence_mt.restype = ctypes.c_int64
    return int(L.orc_linefront_sequence_mt(_p(frames), n, w, h, ctypes.byref(param), int(bool(smoothed)), int(threads)))


def frontend_sequence(frames, num_octaves=1, max_lines=8192, threads=1):
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    return int(lib().orc_frontend_sequence_mt(_p(frames), n, w, h, int(num_octaves), int(max_lines),
                                              int(threads)))


# ---- vanishing-point stage (orc_vp.c / oracle/_ref/libref_vp.so) -------------------
VP_GRID_SHAPE = (90, 360)
_REF_VP_SO = os.path.join(_HERE, "_ref", "libref_vp.so")
_ref_vp = None


class GRand:
    """glibc srand()/rand() restated (orc_grand_*)."""

    class _S(ctypes.Structure):
        _fields_ = [("r", ctypes.c_int32 * 31), ("f", ctypes.c_int), ("b", ctypes.c_int)]

    def __init__(self, seed):
        self.s = GRand._S()
        lib().orc_grand_seed(ctypes.byref(self.s), ctypes.c_uint(seed))

    def next(self):
        return int(lib().orc_grand_next(ctypes.byref(self.s)))


def vp_hypothesis_count():
    return int(lib().orc_vp_hypothesis_count())


def vp_detect(lines, all_lines=None, f=460.0, cx=376.0, cy=240.0, seed=1, frame_count=0, math_mode=1, details=False):
    """vanishing_point_detection::run_vanishing_point_detection -> (vps float64[3,3], vp_idx int32[n_all]);
    details=True adds dict(grid, best_idx, pairs, flags).  math_mode 0 = libm (equals oracle/_ref bit for
    bit), 1 = the shared deterministic functions (what the device computes)."""
    L = lib(); L.orc_vp_detect.restype = ctypes.c_int
    ln = np.ascontiguousarray(lines, LINE_DTYPE)
    al = ln if all_lines is None else np.ascontiguousarray(all_lines, LINE_DTYPE)
    vps = np.zeros((3, 3), np.float64); idx = np.full(len(al), -1, np.int32)
    grid = np.zeros(VP_GRID_SHAPE, np.float64); best = ctypes.c_int32(); flags = ctypes.c_int32()
    pairs = np.zeros((vp_hypothesis_count(), 2), np.int32)
    scores = np.zeros(vp_hypothesis_count() * 360, np.float64) if details else None
    rc = L.orc_vp_detect(_p(ln), len(ln), _p(al), len(al), ctypes.c_float(f), ctypes.c_float(cx), ctypes.c_float(cy),
                         ctypes.c_uint(seed), int(frame_count), int(math_mode), _p(vps), _p(idx), _p(grid),
                         ctypes.byref(best), _p(pairs), ctypes.byref(flags), _p(scores) if scores is not None else None)
    if rc:
        raise ValueError("orc_vp_detect -> %d" % rc)
    if details:
        return vps, idx, dict(grid=grid, best_idx=best.value, pairs=pairs, flags=flags.value, scores=scores)
    return vps, idx


def ref_vp_available():
    return os.path.exists(_REF_VP_SO) or os.path.isdir("/root/reference/feature_tracker/src")


def ref_vp_lib():
    global _ref_vp
    if _ref_vp is None:
        if not os.path.exists(_REF_VP_SO):
            build_ref()
        _ref_vp = ctypes.CDLL(_REF_VP_SO)
        _ref_vp.ref_vp_detect.restype = ctypes.c_int
        _ref_vp.ref_vp_sequence.restype = ctypes.c_longlong
    return _ref_vp


def ref_vp_detect(lines, all_lines=None, f=460.0, cx=376.0, cy=240.0, seed=1, frame_count=0):
    """The same call through the reference's own vanishing_point_detection.cpp (oracle/_ref/libref_vp.so), its
    time(NULL) answered with `seed`."""
    ln = np.ascontiguousarray(lines, LINE_DTYPE)
    al = ln if all_lines is None else np.ascontiguousarray(all_lines, LINE_DTYPE)
    vps = np.zeros((3, 3), np.float64); idx = np.full(len(al), -1, np.int32)
    rc = ref_vp_lib().ref_vp_detect(_p(ln), len(ln), _p(al), len(al), ctypes.c_float(f), ctypes.c_float(cx),
                                    ctypes.c_float(cy), ctypes.c_uint(seed), int(frame_count), _p(vps), _p(idx))
    if rc:
        raise ValueError("ref_vp_detect -> %d" % rc)
    return vps, idx


def vp_sequence(lines, counts, seeds, f=460.0, cx=376.0, cy=240.0, frame_count0=0, math_mode=0, use_ref=False):
    """Frames one after another (timing / sequence parity): lines (n, cap) LINE_DTYPE, counts (n,), seeds (n,)
    -> (vps (n,3,3), vp_idx (n,cap), number of labelled lines)."""
    lines = np.ascontiguousarray(lines, LINE_DTYPE); n, cap = lines.shape
    counts = np.ascontiguousarray(counts, np.int32); seeds = np.ascontiguousarray(seeds, np.uint32)
    vps = np.zeros((n, 3, 3), np.float64); idx = np.full((n, cap), 3, np.int32)
    if use_ref:
        tot = ref_vp_lib().ref_vp_sequence(_p(lines), _p(counts), n, cap, ctypes.c_float(f), ctypes.c_float(cx),
                                           ctypes.c_float(cy), _p(seeds), int(frame_count0), _p(vps), _p(idx))
    else:
        L = lib(); L.orc_vp_sequence.restype = ctypes.c_int64
        tot = L.orc_vp_sequence(_p(lines), _p(counts), n, cap, ctypes.c_float(f), ctypes.c_float(cx), ctypes.c_float(cy),
                                _p(seeds), int(frame_count0), int(math_mode), _p(vps), _p(idx))
    return vps, idx, int(tot)


def line_cloud(lines, ids, line_vps, fx, fy, cx, cy, num_of_cam=1, cam=0):
    """The PointCloud body of line_feature_tracker_node.cpp:64-153 -> dict(points (n,3), id, u, v, vp_x, vp_y, vp_z,
    vp_z_inv) float32; line_vps: (n,4) float64 per-line Vector4d (None / empty = the vp.empty() branch)."""
    ln = np.ascontiguousarray(lines, LINE_DTYPE); n = len(ln)
    ids = np.ascontiguousarray(ids, np.int32)
    lv = None if line_vps is None else np.ascontiguousarray(line_vps, np.float64)
    out = np.zeros(10 * n, np.float32)
    lib().orc_line_cloud(_p(ln), _p(ids), n, _p(lv) if lv is not None else None, 0 if lv is None else len(lv),
                         ctypes.c_float(fx), ctypes.c_float(fy), ctypes.c_float(cx), ctypes.c_float(cy), int(num_of_cam),
                         int(cam), _p(out))
    names = ("id", "u", "v", "vp_x", "vp_y", "vp_z", "vp_z_inv")
    d = {"points": out[:3 * n].reshape(n, 3).copy()}
    for k, nm in enumerate(names):
        d[nm] = out[3 * n + k * n:3 * n + (k + 1) * n].copy()
    return d


# ---- the reference's own LineFeatureTracker::readImage (oracle/_ref/libref_tracker.so) -----------
_REF_TR_SO = os.path.join(_HERE, "_ref", "libref_tracker.so")
_ref_tr = None


def ref_tracker_available():
    return os.path.exists(_REF_TR_SO) or os.path.isdir("/root/reference/feature_tracker/src")


def ref_tracker_lib():
    global _ref_tr
    if _ref_tr is None:
        if not os.path.exists(_REF_TR_SO):
            build_ref()
        _ref_tr = ctypes.CDLL(_REF_TR_SO)
        _ref_tr.ref_tracker_create.restype = ctypes.c_void_p
        _ref_tr.ref_tracker_read.restype = ctypes.c_int
    return _ref_tr


class RefTracker:
    """The reference's LineFeatureTracker (its own line_feature_tracker.cpp), set up as its node's main() does.
    One object per process at a time (the reference reads its parameters from globals)."""

    def __init__(self, mapx, mapy, fx, fy, cx, cy, equalize=True, max_h_lines=25, max_v_lines=25, min_line_length=35.0,
                 line_fit_err=1.8):
        L = ref_tracker_lib()
        self.mapx = np.ascontiguousarray(mapx, np.float32); self.mapy = np.ascontiguousarray(mapy, np.float32)
        h, w = self.mapx.shape
        self.h = L.ref_tracker_create(_p(self.mapx), _p(self.mapy), w, h, ctypes.c_float(fx), ctypes.c_float(fy),
                                      ctypes.c_float(cx), ctypes.c_float(cy), int(bool(equalize)), int(max_h_lines),
                                      int(max_v_lines), ctypes.c_float(min_line_length), ctypes.c_float(line_fit_err))
        self.h = ctypes.c_void_p(self.h)

    def close(self):
        if self.h:
            ref_tracker_lib().ref_tracker_destroy(self.h)
            self.h = None

    def read(self, raw, seed, cap=4096):
        raw = _u8(raw); hh, w = raw.shape
        lines = np.zeros(cap, LINE_DTYPE); ids = np.zeros(cap, np.int32); vps = np.zeros((cap, 4), np.float64)
        t_cnt = np.zeros(cap, np.int32)
        n_vps = ctypes.c_int32(); n_t = ctypes.c_int32(); ex = ctypes.c_int32()
        n = ref_tracker_lib().ref_tracker_read(self.h, _p(raw), w, hh, ctypes.c_uint(seed), cap, _p(lines), _p(ids), _p(vps),
                                               ctypes.byref(n_vps), _p(t_cnt), ctypes.byref(n_t), ctypes.byref(ex))
        if n < 0:
            raise ValueError("ref_tracker_read: cap %d too small" % cap)
        return dict(lines=lines[:n].copy(), ids=ids[:n].tolist(), vps=vps[:n_vps.value].copy(),
                    t_cnt=t_cnt[:n_t.value].tolist(), lines_exit=bool(ex.value))
