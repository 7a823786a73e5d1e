"""ctypes binding of libvplines_b200.so (the C ABI in include/vpl_capi.h).

This is the thin host-side shim the tests and bench.py call through; it adds no
computation.  Loading fails loudly when the library has not been built, and
Context() fails loudly when there is no CUDA device: there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvplines_b200.so")

VPL_OK, VPL_E_INVALID, VPL_E_CUDA, VPL_E_CAPACITY, VPL_E_NODEVICE = 0, -1, -2, -3, -4
STAGES = ["h2d", "pyramid", "scale", "angle", "order", "region", "nfa", "pack", "lbd", "match", "d2h", "preproc",
          "ed_grad", "ed_anchor", "ed_walk", "ed_fit", "lm_pyramid", "lm_track", "lm_vote", "vp_prep", "vp_vote", "vp_score",
          "vp_classify"]

KEYLINE_DTYPE = np.dtype(
    [("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"), ("pt_y", "<f4"),
     ("response", "<f4"), ("size", "<f4"), ("startPointX", "<f4"), ("startPointY", "<f4"),
     ("endPointX", "<f4"), ("endPointY", "<f4"), ("sPointInOctaveX", "<f4"),
     ("sPointInOctaveY", "<f4"), ("ePointInOctaveX", "<f4"), ("ePointInOctaveY", "<f4"),
     ("lineLength", "<f4"), ("numOfPixels", "<i4")])
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
SEGMENT_DTYPE = np.dtype([("x1", "<f4"), ("y1", "<f4"), ("x2", "<f4"), ("y2", "<f4"),
                          ("width", "<f8"), ("prec", "<f8"), ("nfa", "<f8")])
# struct Line's numeric fields (line_matching/src/line.h:8-12) / VplLine
LINE_DTYPE = np.dtype([("endpoint", "<f4", 4), ("equation", "<f8", 3), ("center", "<f4", 2),
                       ("length", "<f4"), ("reserved", "<i4")])
assert KEYLINE_DTYPE.itemsize == 68 and DMATCH_DTYPE.itemsize == 16 and SEGMENT_DTYPE.itemsize == 40
assert LINE_DTYPE.itemsize == 56


class EDLineParam(C.Structure):
    """EDLineParam (line_matching/src/edline_detector.h:32-40) / VplEDLineParam; defaults are the
    tracker node's (feature_tracker/src/line_feature_tracker_node.cpp:203, EuRoC yaml)."""
    _fields_ = [("ksize", C.c_int32), ("sigma", C.c_float), ("gradientThreshold", C.c_float),
                ("anchorThreshold", C.c_float), ("scanIntervals", C.c_int32), ("minLineLen", C.c_int32),
                ("lineFitErrThreshold", C.c_double)]

    def __init__(self, ksize=5, sigma=1.0, gradientThreshold=30, anchorThreshold=5, scanIntervals=2,
                 minLineLen=35, lineFitErrThreshold=1.8):
        super().__init__(ksize, sigma, gradientThreshold, anchorThreshold, scanIntervals, minLineLen,
                         lineFitErrThreshold)


class VplConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_width", C.c_int32), ("max_height", C.c_int32),
                ("max_octaves", C.c_int32), ("max_lines", C.c_int32), ("max_batch", C.c_int32),
                ("num_slots", C.c_int32), ("blur_first", C.c_int32), ("profile", C.c_int32),
                ("lsd_path", C.c_int32)]


class LineMatchParam(C.Structure):
    """VplLineMatchParam: LineMatching / KLT / TopologicalFilter defaults of the reference
    (line_matching/src/line_matching.h:14-18, :45-47; line_matching.cpp:14, :630-631)."""
    _fields_ = [("step", C.c_int32), ("closest_line_threshold", C.c_float), ("line_matching_ratio", C.c_float),
                ("line_distance_error_ratio", C.c_float), ("klt_error_threshold", C.c_float),
                ("max_level", C.c_int32), ("max_count", C.c_int32), ("epsilon", C.c_double), ("min_eig", C.c_float),
                ("topo_distance_threshold", C.c_float), ("topo_length_ratio", C.c_float),
                ("topo_violation_ratio", C.c_float), ("illumination_adapt", C.c_int32),
                ("topological_filter", C.c_int32), ("max_anchors", C.c_int32)]

    def __init__(self, **kw):
        super().__init__()
        load().vpl_linematch_default_param(C.byref(self))
        for k, v in kw.items():
            setattr(self, k, v)


EXPORTS = [
    "vpl_default_config", "vpl_create", "vpl_destroy", "vpl_last_error", "vpl_version", "vpl_device_count",
    "vpl_lsd_detect_batch", "vpl_lbd_compute_batch", "vpl_lbd_compute_float_batch", "vpl_match_batch", "vpl_frontend_batch",
    "vpl_frontend_submit", "vpl_frontend_upload", "vpl_frontend_submit_group", "vpl_frontend_run_resident_group", "vpl_frontend_collect", "vpl_frontend_run_resident", "vpl_sync", "vpl_lsd_raw",
    "vpl_host_register", "vpl_host_unregister", "vpl_last_d2h_bytes", "vpl_frontend_collect_dense",
    "vpl_set_preprocess", "vpl_preprocess_batch",
    "vpl_edlines_default_param", "vpl_edlines_configure", "vpl_edlines_detect_batch", "vpl_edlines_submit",
    "vpl_edlines_collect", "vpl_edlines_run_resident", "vpl_debug_edge_chains",
    "vpl_linematch_default_param", "vpl_linematch_configure", "vpl_linematch_batch", "vpl_debug_linematch_points",
    "vpl_linefront_batch", "vpl_linefront_submit", "vpl_linefront_collect", "vpl_linefront_run_resident",
    "vpl_vp_configure", "vpl_vp_detect_batch", "vpl_vp_submit", "vpl_vp_collect", "vpl_vp_run_resident", "vpl_debug_vp", "vpl_vp_pack_cloud", "vpl_match_run_resident", "vpl_debug_popc_peak", "vpl_debug_vp_scores",
    "vpl_readimage_submit", "vpl_readimage_collect", "vpl_readimage_run_resident",
    "vpl_debug_stage", "vpl_debug_candidates", "vpl_get_stage_times", "vpl_reset_stage_times", "vpl_debug_mark", "vpl_debug_timeline", "vpl_debug_nfa_stats", "vpl_set_profile", "vpl_debug_set_engine_ring_cap", "vpl_debug_set_engine",
    "vpl_kernel_launches",
]

_lib = None


class VplError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vpl error {code}: {msg}")
        self.code = code


def load():
    """dlopen the library (no CUDA call is made until a context is created)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    L.vpl_default_config.argtypes = [C.POINTER(VplConfig)]
    L.vpl_default_config.restype = None
    L.vpl_create.argtypes = [C.POINTER(VplConfig), C.POINTER(vp)]
    L.vpl_destroy.argtypes = [vp]
    L.vpl_destroy.restype = None
    L.vpl_last_error.argtypes = [vp]
    L.vpl_last_error.restype = C.c_char_p
    L.vpl_version.restype = C.c_char_p
    L.vpl_lsd_detect_batch.argtypes = [vp, vp, i32, i32, i32, sz, i32, i32, vp, vp, i32]
    L.vpl_lbd_compute_batch.argtypes = [vp, vp, i32, i32, i32, sz, vp, vp, i32, vp]
    L.vpl_lbd_compute_float_batch.argtypes = [vp, vp, i32, i32, i32, sz, vp, vp, i32, vp]
    L.vpl_match_batch.argtypes = [vp, vp, vp, i32, vp, vp, i32, i32, i32, vp]
    L.vpl_frontend_batch.argtypes = [vp, vp, i32, i32, i32, sz, i32, i32, i32, i32, vp, vp, i32, vp, vp]
    L.vpl_frontend_submit.argtypes = [vp, i32, vp, i32, i32, i32, sz, i32, i32, i32, i32]
    L.vpl_frontend_upload.argtypes = [vp, i32, vp, i32, i32, i32, sz]
    L.vpl_frontend_submit_group.argtypes = [vp, i32, vp, vp, i32, i32, i32, i32, i32, vp]
    L.vpl_frontend_run_resident_group.argtypes = [vp, i32, vp, i32]
    L.vpl_frontend_collect.argtypes = [vp, i32, vp, vp, i32, vp, vp]
    L.vpl_frontend_run_resident.argtypes = [vp, i32, i32]
    L.vpl_sync.argtypes = [vp]
    L.vpl_frontend_collect_dense.argtypes = [vp, i32, vp, vp, vp, vp, C.c_int64, vp]
    L.vpl_set_preprocess.argtypes = [vp, vp, vp, i32, i32, C.c_double, i32]
    L.vpl_preprocess_batch.argtypes = [vp, vp, i32, i32, i32, sz, vp]
    L.vpl_host_register.argtypes = [vp, vp, sz]
    L.vpl_host_unregister.argtypes = [vp, vp]
    L.vpl_last_d2h_bytes.argtypes = [vp, i32]
    L.vpl_last_d2h_bytes.restype = C.c_int64
    L.vpl_lsd_raw.argtypes = [vp, vp, i32, i32, sz, vp, vp, i32]
    L.vpl_debug_stage.argtypes = [vp, i32, vp, i32, i32, sz, vp, sz, vp, vp]
    L.vpl_debug_candidates.argtypes = [vp, vp, vp, i32]
    L.vpl_get_stage_times.argtypes = [vp, vp, vp]
    L.vpl_reset_stage_times.argtypes = [vp]
    L.vpl_debug_mark.argtypes = [vp]
    L.vpl_debug_nfa_stats.argtypes = [vp, vp]
    L.vpl_debug_timeline.argtypes = [vp, i32, vp, vp]
    L.vpl_set_profile.argtypes = [vp, i32]
    L.vpl_debug_set_engine_ring_cap.argtypes = [vp, i32]
    L.vpl_debug_set_engine.argtypes = [vp, i32]
    L.vpl_kernel_launches.argtypes = [vp]
    L.vpl_kernel_launches.restype = C.c_int64
    L.vpl_edlines_default_param.argtypes = [C.POINTER(EDLineParam)]
    L.vpl_edlines_default_param.restype = None
    L.vpl_edlines_configure.argtypes = [vp, C.POINTER(EDLineParam)]
    L.vpl_edlines_detect_batch.argtypes = [vp, vp, i32, i32, i32, sz, i32, vp, vp, i32, vp]
    L.vpl_edlines_submit.argtypes = [vp, i32, vp, i32, i32, i32, sz, i32]
    L.vpl_edlines_collect.argtypes = [vp, i32, vp, vp, i32, vp]
    L.vpl_edlines_run_resident.argtypes = [vp, i32]
    L.vpl_debug_edge_chains.argtypes = [vp, i32, vp, i32, vp, i32, vp, vp]
    L.vpl_linematch_default_param.argtypes = [vp]
    L.vpl_linematch_default_param.restype = None
    L.vpl_linematch_configure.argtypes = [vp, vp]
    L.vpl_linematch_batch.argtypes = [vp, vp, vp, i32, i32, i32, sz, vp, vp, vp, vp, i32, vp]
    L.vpl_debug_linematch_points.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, vp]
    L.vpl_linefront_batch.argtypes = [vp, vp, i32, i32, i32, sz, i32, vp, vp, i32, vp]
    L.vpl_linefront_submit.argtypes = [vp, i32, vp, i32, i32, i32, sz, i32]
    L.vpl_linefront_collect.argtypes = [vp, i32, vp, vp, i32, vp]
    L.vpl_linefront_run_resident.argtypes = [vp, i32]
    L.vpl_vp_configure.argtypes = [vp, C.c_float, C.c_float, C.c_float]
    L.vpl_vp_detect_batch.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, i32, vp, vp, vp, vp]
    L.vpl_vp_submit.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, vp, i32]
    L.vpl_vp_collect.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    L.vpl_vp_run_resident.argtypes = [vp, i32]
    L.vpl_debug_vp.argtypes = [vp, i32, vp, vp, vp]
    L.vpl_debug_vp_scores.argtypes = [vp, i32, vp]
    L.vpl_readimage_submit.argtypes = [vp, i32, vp, i32, i32, i32, sz, i32, vp, i32]
    L.vpl_readimage_collect.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, vp]
    L.vpl_readimage_run_resident.argtypes = [vp, i32]
    L.vpl_match_run_resident.argtypes = [vp, i32]
    L.vpl_debug_popc_peak.argtypes = [vp, vp]
    L.vpl_vp_pack_cloud.argtypes = [vp, i32, vp, i32, C.c_float, C.c_float, C.c_float, C.c_float, i32, i32, vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _img_ptrs(frames):
    """frames: (n,h,w) uint8 C-contiguous array or list of 2-D uint8 arrays -> (array of pointers, keepalive, n, w, h, stride)."""
    if isinstance(frames, np.ndarray) and frames.ndim == 3:
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape
        base = frames.ctypes.data
        ptrs = (C.c_void_p * n)(*[base + i * h * w for i in range(n)])
        return ptrs, frames, n, w, h, w
    arrs = [np.ascontiguousarray(f, np.uint8) for f in frames]
    n = len(arrs)
    h, w = arrs[0].shape
    assert all(a.shape == (h, w) for a in arrs)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    return ptrs, arrs, n, w, h, w


class Context:
    """One GPU, `num_slots` batches in flight.  Mirrors VplContext."""

    def __init__(self, device=0, max_width=752, max_height=480, max_octaves=1, max_lines=2048, max_batch=64,
                 num_slots=2, blur_first=True, profile=False, lsd_path=True):
        L = load()
        cfg = VplConfig()
        L.vpl_default_config(C.byref(cfg))
        cfg.device, cfg.max_width, cfg.max_height = device, max_width, max_height
        cfg.max_octaves, cfg.max_lines, cfg.max_batch = max_octaves, max_lines, max_batch
        cfg.num_slots, cfg.blur_first, cfg.profile = num_slots, int(blur_first), int(profile)
        cfg.lsd_path = int(lsd_path)
        self.cfg = cfg
        self._L = L
        h = C.c_void_p()
        r = L.vpl_create(C.byref(cfg), C.byref(h))
        if r != 0:
            raise VplError(r, L.vpl_last_error(None).decode())
        self._h = h
        self.max_lines = max_lines
        self.max_batch = max_batch
        self.num_slots = num_slots

    def close(self):
        if getattr(self, "_h", None):
            self._L.vpl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, r):
        if r != 0:
            raise VplError(r, self._L.vpl_last_error(self._h).decode())

    # -- LSDDetector::detect -------------------------------------------------
    def lsd_detect_batch(self, frames, scale=2, num_octaves=1, cap=None):
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        cap = cap or self.max_lines
        kl = np.zeros((n, cap), KEYLINE_DTYPE)
        counts = np.zeros(n, np.int32)
        self._ck(self._L.vpl_lsd_detect_batch(self._h, ptrs, n, w, h, stride, scale, num_octaves, _ptr(kl),
                                              _ptr(counts), cap))
        return [kl[f, :counts[f]].copy() for f in range(n)]

    # -- BinaryDescriptor::compute --------------------------------------------
    def lbd_compute_batch(self, frames, keylines, cap=None):
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        assert len(keylines) == n
        cap = cap or max(1, max(len(k) for k in keylines))
        kl = np.zeros((n, cap), KEYLINE_DTYPE)
        counts = np.array([len(k) for k in keylines], np.int32)
        for f, k in enumerate(keylines):
            kl[f, :len(k)] = k
        desc = np.zeros((n, cap, 32), np.uint8)
        self._ck(self._L.vpl_lbd_compute_batch(self._h, ptrs, n, w, h, stride, _ptr(kl), _ptr(counts), cap,
                                               _ptr(desc)))
        return [desc[f, :counts[f]].copy() for f in range(n)]

    def lbd_compute_float_batch(self, frames, keylines, cap=None):
        """BinaryDescriptor::compute(..., returnFloatDescr=True): [n_lines, 72] float32 per frame."""
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        cap = cap or max(1, max(len(k) for k in keylines))
        kl = np.zeros((n, cap), KEYLINE_DTYPE)
        counts = np.array([len(k) for k in keylines], np.int32)
        for f, k in enumerate(keylines):
            kl[f, :len(k)] = k
        fd = np.zeros((n, cap, 72), np.float32)
        self._ck(self._L.vpl_lbd_compute_float_batch(self._h, ptrs, n, w, h, stride, _ptr(kl), _ptr(counts), cap,
                                                     _ptr(fd)))
        return [fd[f, :counts[f]].copy() for f in range(n)]

    # -- BinaryDescriptorMatcher::match / knnMatch ------------------------------
    def match_batch(self, queries, trains, k=1):
        n = len(queries)
        assert len(trains) == n
        cap_q = max(1, max(len(q) for q in queries))
        cap_t = max(1, max(len(t) for t in trains))
        q = np.zeros((n, cap_q, 32), np.uint8)
        t = np.zeros((n, cap_t, 32), np.uint8)
        nq = np.array([len(x) for x in queries], np.int32)
        nt = np.array([len(x) for x in trains], np.int32)
        for i in range(n):
            q[i, :nq[i]] = np.asarray(queries[i], np.uint8).reshape(-1, 32)
            t[i, :nt[i]] = np.asarray(trains[i], np.uint8).reshape(-1, 32)
        out = np.zeros((n, cap_q, k), DMATCH_DTYPE)
        self._ck(self._L.vpl_match_batch(self._h, _ptr(q), _ptr(nq), cap_q, _ptr(t), _ptr(nt), cap_t, n, k,
                                         _ptr(out)))
        return [out[i, :nq[i]].copy() for i in range(n)]

    # -- fused front end ---------------------------------------------------------
    def match_batch_into(self, q, nq, t, nt, k, out):
        """The C call as it is: q (n, cap_q, 32) u8, nq (n,) i32, t (n, cap_t, 32), nt, out (n, cap_q, k) DMATCH_DTYPE."""
        self._ck(self._L.vpl_match_batch(self._h, _ptr(q), _ptr(nq), q.shape[1], _ptr(t), _ptr(nt), t.shape[1], q.shape[0], k,
                                         _ptr(out)))

    def match_run_resident(self, k=1):
        self._ck(self._L.vpl_match_run_resident(self._h, int(k)))

    def popc_peak(self):
        """Measured POPC throughput of the device, popc32 per second."""
        v = C.c_double(0.0)
        self._ck(self._L.vpl_debug_popc_peak(self._h, C.byref(v)))
        return v.value

    def frontend_batch(self, frames, scale=2, num_octaves=1, k=1, chain=False, cap=None):
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        cap = cap or self.max_lines
        kl = np.zeros((n, cap), KEYLINE_DTYPE)
        counts = np.zeros(n, np.int32)
        desc = np.zeros((n, cap, 32), np.uint8)
        m = np.zeros((n, cap, max(k, 1)), DMATCH_DTYPE)
        self._ck(self._L.vpl_frontend_batch(self._h, ptrs, n, w, h, stride, scale, num_octaves, k, int(chain),
                                            _ptr(kl), _ptr(counts), cap, _ptr(desc), _ptr(m)))
        return ([kl[f, :counts[f]].copy() for f in range(n)], [desc[f, :counts[f]].copy() for f in range(n)],
                [m[f, :counts[f]].copy() for f in range(n)])

    def submit(self, slot, frames, scale=2, num_octaves=1, k=1, chain=False):
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        self._ck(self._L.vpl_frontend_submit(self._h, slot, ptrs, n, w, h, stride, scale, num_octaves, k,
                                             int(chain)))
        return n

    def upload(self, slot, frames):
        """Upload the slot's NEXT batch while its current one is in flight (vpl_frontend_upload); frames: a
        contiguous (n, h, w) array inside a host_register range."""
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        self._ck(self._L.vpl_frontend_upload(self._h, slot, ptrs, n, w, h, stride))
        return n

    def submit_uploaded(self, slot, n, w, h, scale=2, num_octaves=1, k=1, chain=False):
        """vpl_frontend_submit with imgs == NULL: run the batch that `upload` staged on this slot."""
        self._ck(self._L.vpl_frontend_submit(self._h, slot, None, n, w, h, w, scale, num_octaves, k, int(chain)))
        return n

    def submit_group(self, slots, ns, w, h, scale=2, num_octaves=1, k=1, chain=None):
        """vpl_frontend_submit_group: the uploaded batches of `slots` (ns[i] frames each) enqueued together, their
        region-engine launches behind a barrier across the group."""
        sl = np.ascontiguousarray(slots, np.int32)
        nn = np.ascontiguousarray(ns, np.int32)
        ch = np.ascontiguousarray([1] * len(sl) if chain is None else [int(bool(x)) for x in chain], np.int32)
        self._ck(self._L.vpl_frontend_submit_group(self._h, len(sl), _ptr(sl), _ptr(nn), w, h, scale, num_octaves, k, _ptr(ch)))

    def run_resident_group(self, slots, k=1):
        sl = np.ascontiguousarray(slots, np.int32)
        self._ck(self._L.vpl_frontend_run_resident_group(self._h, len(sl), _ptr(sl), int(k)))

    def collect_into(self, slot, kl, counts, cap, desc, matches):
        self._ck(self._L.vpl_frontend_collect(self._h, slot, _ptr(kl), _ptr(counts), cap, _ptr(desc),
                                              _ptr(matches)))

    def collect_dense_into(self, slot, counts, kl, desc, matches):
        """Dense collect: rows of all frames packed back to back; returns the total row count."""
        total = C.c_int64(0)
        self._ck(self._L.vpl_frontend_collect_dense(self._h, slot, _ptr(counts), _ptr(kl), _ptr(desc), _ptr(matches),
                                                    len(kl), C.byref(total)))
        return int(total.value)

    def set_preprocess(self, mapx=None, mapy=None, size=None, clahe_clip=0.0, clahe_tiles=8):
        """Undistortion maps (float32 HxW each, or None) and CLAHE clip limit (<= 0: off)."""
        if mapx is not None:
            mapx = np.ascontiguousarray(mapx, np.float32); mapy = np.ascontiguousarray(mapy, np.float32)
            h, w = mapx.shape
        else:
            w, h = size
        self._keep_maps = (mapx, mapy)
        self._ck(self._L.vpl_set_preprocess(self._h, _ptr(mapx), _ptr(mapy), w, h, float(clahe_clip), clahe_tiles))

    def preprocess_batch(self, frames):
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        out = np.zeros((n, h, w), np.uint8)
        self._ck(self._L.vpl_preprocess_batch(self._h, ptrs, n, w, h, stride, _ptr(out)))
        return out

    def host_register(self, arr):
        """Pin a numpy frame buffer so that submits from it skip the staging copy."""
        self._ck(self._L.vpl_host_register(self._h, _ptr(arr), arr.nbytes))

    def host_unregister(self, arr):
        self._ck(self._L.vpl_host_unregister(self._h, _ptr(arr)))

    def last_d2h_bytes(self, slot):
        return int(self._L.vpl_last_d2h_bytes(self._h, slot))

    def run_resident(self, slot, k=1):
        self._ck(self._L.vpl_frontend_run_resident(self._h, slot, k))

    def sync(self):
        self._ck(self._L.vpl_sync(self._h))

    # -- EDLineDetector::EDline (the detector the reference really runs) ----------------
    def edlines_configure(self, param=None):
        self._edp = param or EDLineParam()
        self._ck(self._L.vpl_edlines_configure(self._h, C.byref(self._edp)))

    def edlines_detect_batch(self, frames, smoothed=True, cap=None, with_status=False):
        """-> list of LINE_DTYPE arrays (one per frame, (chain, position) order) [, status per frame]."""
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        cap = cap or self.max_lines
        lines = np.zeros((n, cap), LINE_DTYPE)
        counts = np.zeros(n, np.int32)
        status = np.zeros(n, np.int32)
        self._ck(self._L.vpl_edlines_detect_batch(self._h, ptrs, n, w, h, stride, int(bool(smoothed)), _ptr(lines),
                                                  _ptr(counts), cap, _ptr(status)))
        out = [lines[f, :counts[f]].copy() for f in range(n)]
        return (out, status) if with_status else out

    def edlines_submit(self, slot, frames, smoothed=True):
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        self._ck(self._L.vpl_edlines_submit(self._h, slot, ptrs, n, w, h, stride, int(bool(smoothed))))
        return n

    def edlines_collect_into(self, slot, lines, counts, cap, status=None):
        self._ck(self._L.vpl_edlines_collect(self._h, slot, _ptr(lines), _ptr(counts), cap, _ptr(status)))

    def edlines_run_resident(self, slot):
        self._ck(self._L.vpl_edlines_run_resident(self._h, slot))

    def edge_chains(self, frame, w, h):
        """Edge chains of `frame` of the last EDLines batch on slot 0 -> (xy u32 = x | y << 16, sid)."""
        xy = np.zeros(2 * (w * h // 5) + 16, np.uint32)
        sid = np.zeros(w * h // 100 + 2, np.uint32)
        npx, nch = C.c_int32(0), C.c_int32(0)
        self._ck(self._L.vpl_debug_edge_chains(self._h, frame, _ptr(xy), len(xy), _ptr(sid), len(sid) - 1,
                                               C.byref(npx), C.byref(nch)))
        return xy[:npx.value].copy(), sid[:nch.value + 1].copy()

    # -- LineMatching::Matching (the matcher the reference really runs) -------------------
    def linematch_configure(self, param=None):
        self._lmp = param or LineMatchParam()
        self._ck(self._L.vpl_linematch_configure(self._h, C.byref(self._lmp)))

    def linematch_batch(self, imgs_ref, imgs_cur, lines_ref, lines_cur):
        """n independent pairs -> list of int32 arrays ref_to_cur (one per pair)."""
        pr, keep_r, n, w, h, stride = _img_ptrs(imgs_ref)
        pc, keep_c, n2, w2, h2, _ = _img_ptrs(imgs_cur)
        assert n == n2 == len(lines_ref) == len(lines_cur) and (w, h) == (w2, h2)
        cap = max(1, max(max(len(a), len(b)) for a, b in zip(lines_ref, lines_cur)))
        lr = np.zeros((n, cap), LINE_DTYPE); lc = np.zeros((n, cap), LINE_DTYPE)
        nr = np.array([len(a) for a in lines_ref], np.int32); nc = np.array([len(a) for a in lines_cur], np.int32)
        for p in range(n):
            lr[p, :nr[p]] = np.asarray(lines_ref[p]).view(LINE_DTYPE).reshape(-1)
            lc[p, :nc[p]] = np.asarray(lines_cur[p]).view(LINE_DTYPE).reshape(-1)
        out = np.full((n, cap), -1, np.int32)
        self._ck(self._L.vpl_linematch_batch(self._h, pr, pc, n, w, h, stride, _ptr(lr), _ptr(nr), _ptr(lc), _ptr(nc),
                                             cap, _ptr(out)))
        return [out[p, :nr[p]].copy() for p in range(n)]

    def linematch_points(self, pair, cap=1 << 15):
        """Per-anchor results of `pair` of the last match on slot 0."""
        kr = np.zeros((cap, 2), np.float32); kc = np.zeros((cap, 2), np.float32)
        st = np.zeros(cap, np.uint8); er = np.zeros(cap, np.float32); k2l = np.zeros(cap, np.int32)
        n = C.c_int32(0)
        self._ck(self._L.vpl_debug_linematch_points(self._h, pair, _ptr(kr), _ptr(kc), _ptr(st), _ptr(er), _ptr(k2l),
                                                    cap, C.byref(n)))
        n = n.value
        return dict(kps_ref=kr[:n].copy(), kps_cur=kc[:n].copy(), status=st[:n].copy(), err=er[:n].copy(),
                    kp2line=k2l[:n].copy())

    # -- fused: EDline on every frame + Matching(frame f-1, frame f) ----------------------
    def linefront_batch(self, frames, smoothed=True, cap=None):
        """-> (lines per frame, prev_to_cur per frame; entry 0 is empty)."""
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        cap = cap or self.max_lines
        lines = np.zeros((n, cap), LINE_DTYPE)
        counts = np.zeros(n, np.int32)
        p2c = np.full((n, cap), -1, np.int32)
        self._ck(self._L.vpl_linefront_batch(self._h, ptrs, n, w, h, stride, int(bool(smoothed)), _ptr(lines),
                                             _ptr(counts), cap, _ptr(p2c)))
        return ([lines[f, :counts[f]].copy() for f in range(n)],
                [p2c[f, :counts[f - 1]].copy() if f else p2c[0, :0].copy() for f in range(n)])

    def linefront_submit(self, slot, frames, smoothed=True):
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        self._ck(self._L.vpl_linefront_submit(self._h, slot, ptrs, n, w, h, stride, int(bool(smoothed))))
        return n

    def linefront_collect_into(self, slot, lines, counts, cap, prev_to_cur):
        self._ck(self._L.vpl_linefront_collect(self._h, slot, _ptr(lines), _ptr(counts), cap, _ptr(prev_to_cur)))

    def linefront_run_resident(self, slot):
        self._ck(self._L.vpl_linefront_run_resident(self._h, slot))

    # -- vanishing points (vanishing_point_detection::run_vanishing_point_detection) ----------
    def vp_configure(self, f, cx, cy):
        """= vanishing_point_detection::init(f, cx, cy, .)."""
        self._ck(self._L.vpl_vp_configure(self._h, float(f), float(cx), float(cy)))

    @staticmethod
    def _vp_pack(sets):
        cap = max(1, max(len(a) for a in sets))
        arr = np.zeros((len(sets), cap), LINE_DTYPE)
        cnt = np.array([len(a) for a in sets], np.int32)
        for i, a in enumerate(sets):
            arr[i, :cnt[i]] = np.asarray(a).view(LINE_DTYPE).reshape(-1)
        return arr, cnt

    def vp_detect_batch(self, lines, seeds, all_lines=None, frame_count0=0, with_line_vps=False):
        """lines / all_lines: one LINE_DTYPE array per frame (all_lines None = the same sets); seeds: what
        time(NULL) returned per frame -> (vps (n,3,3) float64, [vp_idx per frame], status (n,) int32
        [, line_vps per frame (k,4)])."""
        n = len(lines)
        la, na = self._vp_pack(list(lines) + (list(all_lines) if all_lines is not None else []))
        cap = la.shape[1]
        seeds = np.ascontiguousarray(seeds, np.uint32)
        assert len(seeds) == n
        vps = np.zeros((n, 3, 3), np.float64); idx = np.full((n, cap), -1, np.int32); st = np.zeros(n, np.int32)
        lv = np.zeros((n, cap, 4), np.float64) if with_line_vps else None
        al = la[n:] if all_lines is not None else None
        nal = na[n:] if all_lines is not None else None
        self._ck(self._L.vpl_vp_detect_batch(self._h, _ptr(la[:n]), _ptr(na[:n]), _ptr(al), _ptr(nal), n, cap,
                                             _ptr(seeds), int(frame_count0), _ptr(vps), _ptr(idx), _ptr(lv), _ptr(st)))
        cnt = nal if all_lines is not None else na[:n]
        out = (vps, [idx[i, :cnt[i]].copy() for i in range(n)], st)
        if with_line_vps:
            out += ([lv[i, :cnt[i]].copy() for i in range(n)],)
        return out

    def vp_submit(self, slot, lines, n_lines, seeds, frame_count0=0):
        """lines (n, cap) LINE_DTYPE, n_lines (n,) int32, seeds (n,) uint32: classified set = the same set."""
        n, cap = lines.shape
        self._ck(self._L.vpl_vp_submit(self._h, slot, _ptr(lines), _ptr(n_lines), None, None, n, cap, _ptr(seeds),
                                       int(frame_count0)))
        return n

    def vp_collect_into(self, slot, cap, vps, vp_idx, status=None, line_vps=None):
        self._ck(self._L.vpl_vp_collect(self._h, slot, cap, _ptr(vps), _ptr(vp_idx), _ptr(line_vps), _ptr(status)))

    def vp_pack_cloud(self, line_ids, fx, fy, cx, cy, num_of_cam=1, cam=0, slot=0):
        """The PointCloud body of line_feature_tracker_node.cpp:64-153 for every frame of the batch last collected from
        `slot`; line_ids: one int array per frame -> list of dict(points (n,3), id, u, v, vp_x, vp_y, vp_z, vp_z_inv)."""
        n = len(line_ids)
        cap = max(1, max(len(a) for a in line_ids))
        ids = np.zeros((n, cap), np.int32)
        for i, a in enumerate(line_ids):
            ids[i, :len(a)] = a
        cloud = np.zeros((n, cap * 10), np.float32)
        self._ck(self._L.vpl_vp_pack_cloud(self._h, slot, _ptr(ids), cap, float(fx), float(fy), float(cx), float(cy),
                                           int(num_of_cam), int(cam), _ptr(cloud)))
        out = []
        for i, a in enumerate(line_ids):
            k = len(a)
            d = {"points": cloud[i, :3 * k].reshape(k, 3).copy()}
            for c_, nm in enumerate(("id", "u", "v", "vp_x", "vp_y", "vp_z", "vp_z_inv")):
                d[nm] = cloud[i, 3 * k + c_ * k:3 * k + (c_ + 1) * k].copy()
            out.append(d)
        return out

    def vp_run_resident(self, slot):
        self._ck(self._L.vpl_vp_run_resident(self._h, slot))

    def vp_debug(self, frame, n_it=105):
        """Stage outputs of `frame` of the last batch on slot 0 -> dict(grid, best_idx, pairs)."""
        grid = np.zeros((90, 360), np.float64); best = C.c_int32(0); pairs = np.zeros((n_it, 2), np.int32)
        self._ck(self._L.vpl_debug_vp(self._h, frame, _ptr(grid), C.byref(best), _ptr(pairs)))
        return dict(grid=grid, best_idx=best.value, pairs=pairs)

    # -- fused: readImage's line pipeline (pre-processing, EDLines, matching, vanishing points) --------------
    def readimage_submit(self, slot, frames, seeds, smoothed=True, frame_count0=0):
        ptrs, keep, n, w, h, stride = _img_ptrs(frames)
        seeds = np.ascontiguousarray(seeds, np.uint32)
        assert len(seeds) == n
        self._ck(self._L.vpl_readimage_submit(self._h, slot, ptrs, n, w, h, stride, int(bool(smoothed)), _ptr(seeds),
                                              int(frame_count0)))
        return n

    def readimage_collect_into(self, slot, lines, counts, cap, prev_to_cur, vps, vp_idx, vp_status=None):
        self._ck(self._L.vpl_readimage_collect(self._h, slot, _ptr(lines), _ptr(counts), cap, _ptr(prev_to_cur), _ptr(vps),
                                               _ptr(vp_idx), _ptr(vp_status)))

    def readimage_run_resident(self, slot):
        self._ck(self._L.vpl_readimage_run_resident(self._h, slot))

    def readimage_batch(self, frames, seeds, smoothed=True, frame_count0=0, cap=None):
        """-> (lines per frame, prev_to_cur per frame, vps (n,3,3), labels per frame, status (n,))."""
        n = self.readimage_submit(0, frames, seeds, smoothed, frame_count0)
        cap = cap or self.max_lines
        lines = np.zeros((n, cap), LINE_DTYPE); counts = np.zeros(n, np.int32); p2c = np.full((n, cap), -1, np.int32)
        vps = np.zeros((n, 3, 3), np.float64); idx = np.full((n, cap), -1, np.int32); st = np.zeros(n, np.int32)
        self.readimage_collect_into(0, lines, counts, cap, p2c, vps, idx, st)
        return ([lines[f, :counts[f]].copy() for f in range(n)],
                [p2c[f, :counts[f - 1]].copy() if f else p2c[0, :0].copy() for f in range(n)],
                vps, [idx[f, :counts[f]].copy() for f in range(n)], st)

    def vp_scores(self, frame, n_it=105):
        """Score of every hypothesis of `frame` of the last batch on slot 0 -> float64 (n_it * 360,)."""
        sc = np.zeros(n_it * 360, np.float64)
        self._ck(self._L.vpl_debug_vp_scores(self._h, frame, _ptr(sc)))
        return sc

    # -- raw stages ----------------------------------------------------------------
    def lsd_raw(self, img, cap=1 << 15):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = np.zeros(cap, SEGMENT_DTYPE)
        cnt = C.c_int32(0)
        self._ck(self._L.vpl_lsd_raw(self._h, _ptr(img), w, h, w, _ptr(out), C.byref(cnt), cap))
        return out[:cnt.value].copy()

    def debug_stage(self, which, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        buf = np.zeros(w * h * 4 + 64, np.uint8)
        ow, oh = C.c_int32(0), C.c_int32(0)
        self._ck(self._L.vpl_debug_stage(self._h, which, _ptr(img), w, h, w, _ptr(buf), buf.nbytes, C.byref(ow),
                                         C.byref(oh)))
        ow, oh = ow.value, oh.value
        if which in (0, 1, 3):
            return buf[:ow * oh].reshape(oh, ow).copy()
        if which == 2:
            g = buf[:ow * oh * 4].view(np.int16).reshape(oh, ow, 2)
            return g[..., 0].copy(), g[..., 1].copy()
        if which == 4:
            return buf[:ow * oh * 4].view(np.float32).reshape(oh, ow).copy()
        return buf[:ow * 4].view(np.int32).copy()

    def debug_candidates(self, cap=1 << 16):
        out = np.zeros((cap, 16), np.float64)
        cnt = C.c_int32(0)
        self._ck(self._L.vpl_debug_candidates(self._h, _ptr(out), C.byref(cnt), cap))
        return out[:min(cnt.value, cap)].copy()

    # -- measurement -----------------------------------------------------------------
    def stage_times(self):
        ms = np.zeros(len(STAGES), np.float64)
        ln = np.zeros(len(STAGES), np.int64)
        self._ck(self._L.vpl_get_stage_times(self._h, _ptr(ms), _ptr(ln)))
        return {s: (float(ms[i]), int(ln[i])) for i, s in enumerate(STAGES)}

    def nfa_stats(self):
        """(decisions, answered by the float32 test, disagreements, 0) of a -DVPL_NFA_CHECK build; zeros otherwise."""
        v = np.zeros(4, np.uint64)
        self._ck(self._L.vpl_debug_nfa_stats(self._h, _ptr(v)))
        return [int(x) for x in v]

    def mark(self):
        self._ck(self._L.vpl_debug_mark(self._h))

    def timeline(self, slot):
        """{stage: (start_ms, end_ms)} of the slot's last batch, relative to the last mark() (profile on)."""
        a = np.zeros(len(STAGES), np.float64)
        b = np.zeros(len(STAGES), np.float64)
        self._ck(self._L.vpl_debug_timeline(self._h, slot, _ptr(a), _ptr(b)))
        return {s: (float(a[i]), float(b[i])) for i, s in enumerate(STAGES) if a[i] >= 0}

    def reset_stage_times(self):
        self._ck(self._L.vpl_reset_stage_times(self._h))

    def set_profile(self, on):
        """Per-stage CUDA-event timing on / off (with it on, a submit waits for the slot's previous batch)."""
        self._ck(self._L.vpl_set_profile(self._h, int(bool(on))))

    def set_engine_ring_cap(self, entries_per_lane):
        """Test hook: ring capacity per lane of the LSD region engine (0 = default)."""
        self._ck(self._L.vpl_debug_set_engine_ring_cap(self._h, int(entries_per_lane)))

    def set_engine(self, kind):
        """LSD region engine: 0 = default (warp-cooperative, sequential seed order), 1 = speculative (32 seeds in flight)."""
        self._ck(self._L.vpl_debug_set_engine(self._h, int(kind)))

    def kernel_launches(self):
        return int(self._L.vpl_kernel_launches(self._h))


def device_count():
    return int(load().vpl_device_count())
