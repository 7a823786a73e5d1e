"""CPU oracle of LineFeatureTracker::readImage's bookkeeping (feature_tracker/src/line_feature_tracker.cpp:56-288).

TEST INFRASTRUCTURE ONLY (importable from tests/, smoke() and bench.py's CPU legs; never from the product).

The pixel work is the oracle's (remap + CLAHE: orc_preproc.c, EDLines: orc_edlines.c, line matching: orc_linematch.c,
vanishing points: orc_vp.c -- each pinned bit for bit against cv2 4.13 / the reference's own sources); what is
restated here, literally and with its quirks, is the sequential host logic between those calls:

  :95-107   ids: the first image numbers its lines 0.., later images start every line at -1, t_cnt 0
  :113-126  matches prev -> cur; `if (mt > 0)` (a match to current line 0 is ignored) and
            `t_cnt[mt] = curframe_->t_cnt[mt] + 1` (indexes the PREVIOUS frame's counters by the CURRENT index;
            past their end the reference reads out of bounds -- the oracle reads 0 there and says so in `notes`)
  :128-155  split into tracked / new; the two `verticalLine` tests for tracked lines can never be true
            (angle < pi/4 && angle > 3pi/4), so only new non-horizontal lines reach verticalLine
  :157-176  new lines: "horizontal" when |angle| in [3.14/4, 3*3.14/4] (the literal 3.14), else "vertical"
  :177-229  quota: top up the tracked set with new h lines to max_h_lines and new v lines to max_v_lines
  :231-277  vanishing points on (verticalLine, vecLine) if verticalLine has more than 2 lines, else on
            (vecLine, vecLine); one Vector4d per line: zeros for label 3, else (vp, z/z); nothing when the frame
            keeps 2 lines or fewer ("no vp lines": zeros)
  :285      curframe_.swap(forwframe_) -- t_cnt is NOT swapped with the selection: it keeps one entry per raw
            detection of its frame
The first image never reaches the matching block (curframe_ is empty), so its vps stay empty and the vanishing-point
object's first call (frame_count 0) happens on the SECOND image.
"""
import math

import numpy as np

from . import oracle as O


def seg_angle(line):
    """LineFeatureTracker::segAngle (:20-25): std::atan2 on float differences converted to double."""
    e = line["endpoint"] if "endpoint" in line.dtype.names else line["line_endpoint"]
    x0, y0, x1, y1 = (np.float32(v) for v in e)
    if x1 > x0:
        return math.atan2(float(np.float32(y1 - y0)), float(np.float32(x1 - x0)))
    return math.atan2(float(np.float32(y0 - y1)), float(np.float32(x0 - x1)))


def _is_h(a):
    return (3.14 / 4.0 <= a <= 3 * 3.14 / 4.0) or (-3 * 3.14 / 4.0 <= a <= -3.14 / 4.0)


class Tracker:
    """One LineFeatureTracker object.  read(img, seed) = readImage(img) with time(NULL) == seed."""

    def __init__(self, mapx, mapy, f, cx, cy, equalize=True, max_h_lines=25, max_v_lines=25, min_line_length=35.0,
                 line_fit_err=1.8, math_mode=1):
        self.mapx, self.mapy = mapx, mapy
        self.cam = (f, cx, cy)           # vpdetect.init(K(0,0), K(0,2), K(1,2), 0.5)
        self.equalize = equalize
        self.max_h, self.max_v = max_h_lines, max_v_lines
        self.param = O.EDLineParam(minLineLen=int(min_line_length), lineFitErrThreshold=float(line_fit_err))
        self.math_mode = math_mode
        self.cur = None                  # dict(img, lines, ids, vps, t_cnt)
        self.allfeature_cnt = 0
        self.vp_calls = 0                # vanishing_point_detection::frame_count
        self.lines_exit = True
        self.notes = []

    def read(self, raw, seed):
        self.lines_exit = True
        img = O.remap_linear(raw, self.mapx, self.mapy)
        if self.equalize:
            img = O.clahe(img, 3.0, 8)
        first = self.cur is None
        if first:
            self.cur = dict(img=img, lines=np.zeros(0, O.LINE_DTYPE), ids=[], vps=np.zeros((0, 4)), t_cnt=[])
        lines = O.edline_detect(img, self.param, True)
        if len(lines) == 0:
            self.lines_exit = False
            return self.cur
        n = len(lines)
        ids = list(range(self.allfeature_cnt, self.allfeature_cnt + n)) if first else [-1] * n
        if first:
            self.allfeature_cnt += n
        t_cnt = [0] * n
        fw = dict(img=img, lines=lines, ids=ids, vps=np.zeros((0, 4)), t_cnt=t_cnt)
        cur = self.cur
        if len(cur["lines"]) > 0:
            p2c = O.line_matching(cur["img"], img, cur["lines"], lines, illum=True, topo=True)
            for k, mt in enumerate(p2c):
                mt = int(mt)
                if mt > 0:
                    ids[mt] = cur["ids"][k]
                    if mt < len(cur["t_cnt"]):
                        t_cnt[mt] = cur["t_cnt"][mt] + 1
                    else:  # the reference reads past the end of the previous frame's counters here
                        t_cnt[mt] = 1
                        self.notes.append(("t_cnt_out_of_range", mt, len(cur["t_cnt"])))
            tracked, tracked_ids, new, new_ids, vertical = [], [], [], [], []
            for i in range(n):
                if ids[i] == -1:
                    ids[i] = self.allfeature_cnt
                    self.allfeature_cnt += 1
                    new.append(i); new_ids.append(ids[i])
                else:
                    tracked.append(i); tracked_ids.append(ids[i])
                    a = seg_angle(lines[i])
                    if (a < 3.14 / 4.0 and a > 3 * 3.14 / 4.0) or (a > -3.14 / 4.0 and a < -3 * 3.14 / 4.0):
                        vertical.append(i)  # unreachable, as in the reference
            h_new, h_ids, v_new, v_ids = [], [], [], []
            for i, lid in zip(new, new_ids):
                if _is_h(seg_angle(lines[i])):
                    h_new.append(i); h_ids.append(lid)
                else:
                    v_new.append(i); v_ids.append(lid); vertical.append(i)
            h_line = sum(1 for i in tracked if _is_h(seg_angle(lines[i])))
            v_line = len(tracked) - h_line
            diff_h, diff_v = self.max_h - h_line, self.max_v - v_line
            if diff_h > 0:
                k = min(diff_h, len(h_new))
                tracked += h_new[:k]; tracked_ids += h_ids[:k]
            if diff_v > 0:
                k = min(diff_v, len(v_new))
                tracked += v_new[:k]; tracked_ids += v_ids[:k]
            sel = lines[tracked] if tracked else lines[:0]
            fw["lines"], fw["ids"] = sel, tracked_ids
            vps4 = np.zeros((len(sel), 4))
            if len(sel) > 2:
                vl = lines[vertical] if len(vertical) > 2 else sel
                vps, idx, d = O.vp_detect(vl, sel, *self.cam, seed=seed, frame_count=self.vp_calls,
                                          math_mode=self.math_mode, details=True)
                self.vp_calls += 1
                if d["flags"] & 1:
                    self.notes.append(("vp_lx_out_of_range", seed))
                for i, lab in enumerate(idx):
                    if lab != 3:
                        v = vps[lab]
                        vps4[i] = (v[0], v[1], v[2], v[2] / v[2])
            fw["vps"] = vps4
        self.cur = fw
        return fw
