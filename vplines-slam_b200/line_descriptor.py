"""Python mirror of the OpenCV-3.4 `cv::line_descriptor` surface the north star
keeps (opencv_contrib 3.4 modules/line_descriptor/include/opencv2/line_descriptor/
descriptor.hpp): LSDDetector::detect, BinaryDescriptor::compute,
BinaryDescriptorMatcher::match / knnMatch on ndarray (cv::Mat) / KeyLine / DMatch.
Same names, argument meaning and error behaviour; every call goes through the C ABI
(include/vpl_capi.h) to the CUDA library -- there is no CPU path.  It stands where
the reference's line tracker calls edline_detect / match_line_match
(/root/reference/feature_tracker/src/line_feature_tracker.cpp:87, :115).
"""
import numpy as np

from . import capi

_ctx = None
_ctx_key = None


def _context(w, h, octaves, lines=4096):
    """Process-wide context, grown on demand (the reference's detector/matcher objects
    also own scratch that is reused across frames, edline_detector.cpp:95-123)."""
    global _ctx, _ctx_key
    if _ctx is not None:
        mw, mh, mo, ml = _ctx_key
        if w <= mw and h <= mh and octaves <= mo and lines <= ml:
            return _ctx
        w, h, octaves, lines = max(w, mw), max(h, mh), max(octaves, mo), max(lines, ml)
        _ctx.close()
    _ctx = capi.Context(max_width=w, max_height=h, max_octaves=octaves, max_lines=lines, max_batch=4, num_slots=2)
    _ctx_key = (w, h, octaves, lines)
    return _ctx


class KeyLine:
    """cv::line_descriptor::KeyLine."""
    __slots__ = ("angle", "class_id", "octave", "pt", "response", "size", "startPointX", "startPointY",
                 "endPointX", "endPointY", "sPointInOctaveX", "sPointInOctaveY", "ePointInOctaveX",
                 "ePointInOctaveY", "lineLength", "numOfPixels")

    def __init__(self):
        self.angle = 0.0; self.class_id = -1; self.octave = 0; self.pt = (0.0, 0.0); self.response = 0.0
        self.size = 0.0; self.startPointX = self.startPointY = self.endPointX = self.endPointY = 0.0
        self.sPointInOctaveX = self.sPointInOctaveY = self.ePointInOctaveX = self.ePointInOctaveY = 0.0
        self.lineLength = 0.0; self.numOfPixels = 0

    def getStartPoint(self):
        return (self.startPointX, self.startPointY)

    def getEndPoint(self):
        return (self.endPointX, self.endPointY)

    def getStartPointInOctave(self):
        return (self.sPointInOctaveX, self.sPointInOctaveY)

    def getEndPointInOctave(self):
        return (self.ePointInOctaveX, self.ePointInOctaveY)


class DMatch:
    """cv::DMatch."""
    __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

    def __init__(self, queryIdx=-1, trainIdx=-1, imgIdx=-1, distance=float("inf")):
        self.queryIdx, self.trainIdx, self.imgIdx, self.distance = queryIdx, trainIdx, imgIdx, distance

    def __lt__(self, o):
        return self.distance < o.distance

    def __repr__(self):
        return f"DMatch(q={self.queryIdx}, t={self.trainIdx}, d={self.distance})"


def keylines_from_records(rec):
    out = []
    for r in rec:
        k = KeyLine()
        k.angle = float(r["angle"]); k.class_id = int(r["class_id"]); k.octave = int(r["octave"])
        k.pt = (float(r["pt_x"]), float(r["pt_y"])); k.response = float(r["response"]); k.size = float(r["size"])
        k.startPointX = float(r["startPointX"]); k.startPointY = float(r["startPointY"])
        k.endPointX = float(r["endPointX"]); k.endPointY = float(r["endPointY"])
        k.sPointInOctaveX = float(r["sPointInOctaveX"]); k.sPointInOctaveY = float(r["sPointInOctaveY"])
        k.ePointInOctaveX = float(r["ePointInOctaveX"]); k.ePointInOctaveY = float(r["ePointInOctaveY"])
        k.lineLength = float(r["lineLength"]); k.numOfPixels = int(r["numOfPixels"])
        out.append(k)
    return out


def records_from_keylines(kls):
    if isinstance(kls, np.ndarray) and kls.dtype == capi.KEYLINE_DTYPE:
        return kls
    rec = np.zeros(len(kls), capi.KEYLINE_DTYPE)
    for i, k in enumerate(kls):
        rec[i] = (k.angle, k.class_id, k.octave, k.pt[0], k.pt[1], k.response, k.size, k.startPointX,
                  k.startPointY, k.endPointX, k.endPointY, k.sPointInOctaveX, k.sPointInOctaveY,
                  k.ePointInOctaveX, k.ePointInOctaveY, k.lineLength, k.numOfPixels)
    return rec


def _gray_u8(image, what):
    image = np.asarray(image)
    if image.dtype != np.uint8:
        # LSDDetector.cpp / binary_descriptor.cpp: throw std::runtime_error( "Error, depth image!= 0" )
        raise RuntimeError("Error, depth image!= 0")
    if image.ndim == 3:
        raise RuntimeError(f"{what}: only CV_8UC1 images are accepted by the B200 path (convert with cvtColor first)")
    return np.ascontiguousarray(image)


class LSDDetector:
    """cv::line_descriptor::LSDDetector."""

    @staticmethod
    def createLSDDetector():
        return LSDDetector()

    def detect(self, image, scale, numOctaves, mask=None, as_records=False):
        """void detect(const Mat& image, std::vector<KeyLine>& keypoints, int scale, int numOctaves, const Mat& mask)"""
        image = _gray_u8(image, "LSDDetector::detect")
        if mask is not None:
            mask = np.asarray(mask)
            if mask.shape != image.shape or mask.dtype != np.uint8:
                # "Mask error while detecting lines: please check its dimensions and that data type is CV_8UC1"
                raise RuntimeError("Mask error while detecting lines: please check its dimensions and that data "
                                   "type is CV_8UC1")
        h, w = image.shape
        ctx = _context(w, h, numOctaves)
        rec = ctx.lsd_detect_batch(image[None], scale=scale, num_octaves=numOctaves)[0]
        if mask is not None and len(rec):
            sy = rec["startPointY"].astype(np.int32); sx = rec["startPointX"].astype(np.int32)
            ey = rec["endPointY"].astype(np.int32); ex = rec["endPointX"].astype(np.int32)
            drop = (mask[sy, sx] == 0) & (mask[ey, ex] == 0)
            rec = rec[~drop]
        return rec if as_records else keylines_from_records(rec)


class BinaryDescriptor:
    """cv::line_descriptor::BinaryDescriptor (compute only; the detector of the path is LSDDetector)."""

    @staticmethod
    def createBinaryDescriptor():
        return BinaryDescriptor()

    def descriptorSize(self):
        return 32

    def compute(self, image, keylines, returnFloatDescr=False):
        """void compute(const Mat& image, std::vector<KeyLine>& keylines, Mat& descriptors, bool returnFloatDescr)"""
        image = _gray_u8(image, "BinaryDescriptor::compute")
        if len(keylines) == 0:
            print("Error: keypoint list is empty")
            return np.zeros((0, 32), np.uint8)
        rec = records_from_keylines(keylines)
        h, w = image.shape
        ctx = _context(w, h, int(rec["octave"].max()) + 1, max(4096, len(rec)))
        if returnFloatDescr:
            return ctx.lbd_compute_float_batch(image[None], [rec])[0]
        return ctx.lbd_compute_batch(image[None], [rec])[0]


class BinaryDescriptorMatcher:
    """cv::line_descriptor::BinaryDescriptorMatcher, brute-force semantics."""

    @staticmethod
    def createBinaryDescriptorMatcher():
        return BinaryDescriptorMatcher()

    @staticmethod
    def _check(q, t):
        q = np.asarray(q); t = np.asarray(t)
        if q.size == 0 or t.size == 0:
            print("Error: descriptors matrices cannot be void")
            return None, None
        if q.dtype != np.uint8 or t.dtype != np.uint8 or q.shape[-1] != 32 or t.shape[-1] != 32:
            raise RuntimeError("descriptors must be CV_8UC1 with 32 columns")
        return np.ascontiguousarray(q).reshape(-1, 32), np.ascontiguousarray(t).reshape(-1, 32)

    def knnMatch(self, queryDescriptors, trainDescriptors, k, mask=None, compactResult=False, as_records=False):
        q, t = self._check(queryDescriptors, trainDescriptors)
        if q is None:
            return []
        ctx = _context(8, 8, 1, max(4096, len(q), len(t)))
        m = ctx.match_batch([q], [t], k=k)[0]
        if as_records:
            return m
        out = []
        for i in range(len(q)):
            if mask is not None and np.asarray(mask).reshape(-1)[i] == 0:
                if not compactResult:
                    out.append([])
                continue
            out.append([DMatch(int(r["queryIdx"]), int(r["trainIdx"]), 0, float(r["distance"]))
                        for r in m[i] if r["trainIdx"] >= 0])
        return out

    def match(self, queryDescriptors, trainDescriptors, mask=None):
        res = self.knnMatch(queryDescriptors, trainDescriptors, 1, mask=mask, compactResult=True)
        return [r[0] for r in res if r]
