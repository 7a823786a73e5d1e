// linefeature_tracker_b200.hpp -- the reference's LineFeatureTracker (feature_tracker/include/linefeature_tracker.h:57-88,
// feature_tracker/src/line_feature_tracker.cpp:56-288), over libvplines_b200.so.
//
// Same class, members and per-frame results as the reference's tracker object: readImage(const cv::Mat&) fills
// curframe_->{img, vecLine, lineID, vps, t_cnt} exactly as the reference does --
//   remap (undistortion, INTER_LINEAR) + CLAHE(3.0, 8x8) if EQUALIZE                    :62-68    device (vpl_preprocess_batch)
//   EDLineDetector::EDline(img, lines, smoothed = true)                                 :87       device
//   ids / t_cnt of a new frame                                                          :95-107   host
//   LineMatching::Matching(prev img, img, prev SELECTED lines, lines)                   :115      device
//   id / track-count propagation, with `if (mt > 0)` and `t_cnt[mt]` as written         :118-126  host
//   tracked / new split, horizontal / vertical quota (max_h_lines, max_v_lines)         :128-230  host
//   vanishing points on (verticalLine | vecLine, vecLine), one Vector4d per line        :233-277  device + host
//   curframe_.swap(forwframe_)                                                          :285
// The host part is the small sequential bookkeeping between the kernels; it is restated literally, quirks included
// (the two verticalLine tests for tracked lines that can never hold, the literal 3.14, t_cnt never swapped with the
// selection).  Pinned: tests/golden/ref_tracker.npz holds what the reference's own line_feature_tracker.cpp produces
// for the bundled EuRoC sequence (compiled into oracle/_ref/libref_tracker.so against oracle/cvshim);
// tests/test_gpu_tracker.py asserts that this class reproduces every line, id and vanishing-point vector of it.
//
// Differences, all forced by what is absent here:
//   - readIntrinsicParameter(calib_file) needs camodocal (Ceres, Eigen): setIntrinsics(mapx, mapy, w, h, fx, fy, cx, cy)
//     takes the undistortion maps and K that Camera::initUndistortRectifyMap returns (:30-33);
//   - EQUALIZE, max_h_lines, max_v_lines, MIN_LINE_LENGTH, line_fit_err are members instead of the globals of
//     feature_tracker/src/parameters.cpp;
//   - where the reference reads curframe_->t_cnt past its end (:122, a heap over-read) the count restarts at 1;
//   - time(NULL) seeds the vanishing-point stage as in the reference; vpdetect.seed_source makes runs repeatable.
// There is no CPU fallback: every call throws std::runtime_error(vpl_last_error()) when the device path fails.
#pragma once
#include <array>
#include <cmath>
#include <memory>
#include <vector>

#include "line_matching_b200.hpp"

namespace vplines {
namespace ref {

struct Vector4d {  // stands in for Eigen::Vector4d (the reference stores one per line, :246-262)
  double v[4];
  Vector4d() : v{0, 0, 0, 0} {}
  Vector4d(double a, double b, double c, double d) : v{a, b, c, d} {}
  double& operator()(int i) { return v[i]; }
  const double& operator()(int i) const { return v[i]; }
};

class FrameLines {  // linefeature_tracker.h:44-55
 public:
  cv::Mat img;
  std::vector<Line> vecLine;
  std::vector<int> lineID;
  std::vector<Vector4d> vps;
  std::vector<int> vp_idx;
  std::vector<int> t_cnt;
};
typedef std::shared_ptr<FrameLines> FrameLinesPtr;

class LineFeatureTracker {
 public:
  LineFeatureTracker() : lines_exit(true), allfeature_cnt(0) {}  // line_feature_tracker.cpp:7-17
  ~LineFeatureTracker() { if (pre_) vpl_destroy(pre_); }
  LineFeatureTracker(const LineFeatureTracker&) = delete;
  LineFeatureTracker& operator=(const LineFeatureTracker&) = delete;

  // what readIntrinsicParameter (:26-34) obtains from the camera model: CV_32FC1 undistortion maps (w x h each) and K
  void setIntrinsics(const float* mapx, const float* mapy, int w, int h, float fx, float fy, float cx, float cy) {
    mapx_.assign(mapx, mapx + (size_t)w * h);
    mapy_.assign(mapy, mapy + (size_t)w * h);
    w_ = w; h_ = h; fx_ = fx; fy_ = fy; cx_ = cx; cy_ = cy;
    vpdetect.init(fx, cx, cy, 0.5);  // K(0,0), K(0,2), K(1,2)
    pre_set_ = false;
  }

  std::vector<Line> undistortedLineEndPoints() {  // :36-52
    std::vector<Line> un_lines = curframe_->vecLine;
    for (size_t i = 0; i < curframe_->vecLine.size(); i++) {
      un_lines[i].line_endpoint[0] = (curframe_->vecLine[i].line_endpoint[0] - cx_) / fx_;
      un_lines[i].line_endpoint[1] = (curframe_->vecLine[i].line_endpoint[1] - cy_) / fy_;
      un_lines[i].line_endpoint[2] = (curframe_->vecLine[i].line_endpoint[2] - cx_) / fx_;
      un_lines[i].line_endpoint[3] = (curframe_->vecLine[i].line_endpoint[3] - cy_) / fy_;
    }
    return un_lines;
  }

  void readImage(const cv::Mat& _img) {  // :56-288
    lines_exit = true;
    cv::Mat img = preprocess(_img);
    bool first_img = false;
    if (forwframe_ == nullptr) {
      forwframe_.reset(new FrameLines);
      curframe_.reset(new FrameLines);
      forwframe_->img = img;
      curframe_->img = img;
      first_img = true;
    } else {
      forwframe_.reset(new FrameLines);
      forwframe_->img = img;
    }
    line_detctor.EDline(forwframe_->img, forwframe_->vecLine, true);
    if (forwframe_->vecLine.empty()) {
      lines_exit = false;
      return;
    }
    for (size_t i = 0; i < forwframe_->vecLine.size(); ++i) {
      forwframe_->lineID.emplace_back(first_img ? allfeature_cnt++ : -1);
      forwframe_->t_cnt.emplace_back(0);
    }
    if (curframe_->vecLine.size() > 0) {
      std::vector<int> line_prev_to_line_cur;
      line_matching.Matching(curframe_->img, forwframe_->img, curframe_->vecLine, forwframe_->vecLine, line_prev_to_line_cur,
                             *(const cv::Mat*)nullptr, *(const cv::Mat*)nullptr, *(const cv::Mat*)nullptr, true, true, 0);
      for (size_t k = 0; k < line_prev_to_line_cur.size(); ++k) {
        int mt = line_prev_to_line_cur[k];
        if (mt > 0) {  // as written: a match to current line 0 is ignored
          forwframe_->lineID[(size_t)mt] = curframe_->lineID[k];
          // as written: the PREVIOUS frame's counter at the CURRENT index (past its end the reference over-reads)
          forwframe_->t_cnt[(size_t)mt] = ((size_t)mt < curframe_->t_cnt.size() ? curframe_->t_cnt[(size_t)mt] : 0) + 1;
        }
      }
      std::vector<Line> vecLine_tracked, vecLine_new, verticalLine;
      std::vector<int> lineID_tracked, lineID_new;
      for (size_t i = 0; i < forwframe_->vecLine.size(); ++i) {
        if (forwframe_->lineID[i] == -1) {
          forwframe_->lineID[i] = allfeature_cnt++;
          vecLine_new.emplace_back(forwframe_->vecLine[i]);
          lineID_new.emplace_back(forwframe_->lineID[i]);
        } else {
          vecLine_tracked.emplace_back(forwframe_->vecLine[i]);
          lineID_tracked.emplace_back(forwframe_->lineID[i]);
          double angle = segAngle(forwframe_->vecLine[i]);
          if ((angle < 3.14 / 4.0 && angle > 3 * 3.14 / 4.0) || (angle > -3.14 / 4.0 && angle < -3 * 3.14 / 4.0))
            verticalLine.push_back(forwframe_->vecLine[i]);  // never true, as in the reference
        }
      }
      std::vector<Line> h_Line_new, v_Line_new;
      std::vector<int> h_lineID_new, v_lineID_new;
      for (size_t i = 0; i < vecLine_new.size(); ++i) {
        double angle = segAngle(vecLine_new[i]);
        if (isHorizontal(angle)) {
          h_Line_new.emplace_back(vecLine_new[i]);
          h_lineID_new.emplace_back(lineID_new[i]);
        } else {
          v_Line_new.emplace_back(vecLine_new[i]);
          v_lineID_new.emplace_back(lineID_new[i]);
          verticalLine.push_back(vecLine_new[i]);
        }
      }
      int h_line = 0, v_line = 0;
      for (size_t i = 0; i < vecLine_tracked.size(); ++i) {
        if (isHorizontal(segAngle(vecLine_tracked[i]))) h_line++;
        else v_line++;
      }
      int diff_h = max_h_lines - h_line;
      int diff_v = max_v_lines - v_line;
      if (diff_h > 0) {
        if ((size_t)diff_h > h_Line_new.size()) diff_h = (int)h_Line_new.size();
        for (int k = 0; k < diff_h; ++k) {
          vecLine_tracked.emplace_back(h_Line_new[(size_t)k]);
          lineID_tracked.emplace_back(h_lineID_new[(size_t)k]);
        }
      }
      if (diff_v > 0) {
        if ((size_t)diff_v > v_Line_new.size()) diff_v = (int)v_Line_new.size();
        for (int k = 0; k < diff_v; ++k) {
          vecLine_tracked.emplace_back(v_Line_new[(size_t)k]);
          lineID_tracked.emplace_back(v_lineID_new[(size_t)k]);
        }
      }
      forwframe_->vecLine.swap(vecLine_tracked);
      forwframe_->lineID.swap(lineID_tracked);
      forwframe_->vps.clear();
      if (forwframe_->vecLine.size() > 2) {
        std::vector<Vector3d> _vps;
        std::vector<int> local_vp_ids;
        if (verticalLine.size() > 2)
          vpdetect.run_vanishing_point_detection(forwframe_->img, verticalLine, forwframe_->vecLine, _vps, local_vp_ids);
        else
          vpdetect.run_vanishing_point_detection(forwframe_->img, forwframe_->vecLine, forwframe_->vecLine, _vps, local_vp_ids);
        for (size_t i = 0; i < forwframe_->vecLine.size(); i++) {
          if (local_vp_ids.size() > 0 && local_vp_ids[i] != 3) {
            const Vector3d& v = _vps[(size_t)local_vp_ids[i]];
            forwframe_->vps.push_back(Vector4d(v.x(), v.y(), v.z(), v.z() / v.z()));
          } else {
            forwframe_->vps.push_back(Vector4d(0.0, 0.0, 0.0, 0.0));
          }
        }
      } else {
        for (size_t i = 0; i < forwframe_->vecLine.size(); i++) forwframe_->vps.push_back(Vector4d(0.0, 0.0, 0.0, 0.0));
      }
    }
    curframe_.swap(forwframe_);
  }

  FrameLinesPtr curframe_, forwframe_;
  vanishing_point_detection vpdetect;
  bool lines_exit;
  LineMatching line_matching;
  EDLineDetector line_detctor;
  // the globals of feature_tracker/src/parameters.cpp this class reads (EuRoC yaml values)
  int EQUALIZE = 1;
  int max_h_lines = 25;
  int max_v_lines = 25;

 private:
  static double segAngle(const Line& s) {  // :20-25
    if (s.line_endpoint[2] > s.line_endpoint[0])
      return std::atan2(s.line_endpoint[3] - s.line_endpoint[1], s.line_endpoint[2] - s.line_endpoint[0]);
    return std::atan2(s.line_endpoint[1] - s.line_endpoint[3], s.line_endpoint[0] - s.line_endpoint[2]);
  }
  static bool isHorizontal(double angle) {
    return (angle >= 3.14 / 4.0 && angle <= 3 * 3.14 / 4.0) || (angle <= -3.14 / 4.0 && angle >= -3 * 3.14 / 4.0);
  }
  // cv::remap(INTER_LINEAR) + CLAHE(3.0, 8x8) on the device, through a small context of this object's own (the
  // detector's context must not carry the maps: it would undistort the frames it is handed a second time)
  cv::Mat preprocess(const cv::Mat& raw) {
    detail::need_u8(raw);
    if (mapx_.empty() || raw.cols != w_ || raw.rows != h_)
      throw std::runtime_error("LineFeatureTracker::readImage: call setIntrinsics with maps of the image size first");
    if (!pre_) {
      VplConfig c;
      vpl_default_config(&c);
      c.max_width = w_; c.max_height = h_; c.max_lines = 16; c.max_batch = 1; c.num_slots = 1; c.lsd_path = 0;
      int r = vpl_create(&c, &pre_);
      if (r != VPL_OK) throw std::runtime_error(std::string("vpl_create: ") + vpl_last_error(nullptr));
      pre_set_ = false;
    }
    if (!pre_set_) {
      detail::RefCtx::check(pre_, vpl_set_preprocess(pre_, mapx_.data(), mapy_.data(), w_, h_, EQUALIZE ? 3.0 : 0.0, 8),
                            "vpl_set_preprocess");
      pre_set_ = true;
    }
    cv::Mat out(h_, w_, CV_8UC1);
    const uint8_t* ptr = raw.data;
    detail::RefCtx::check(pre_, vpl_preprocess_batch(pre_, &ptr, 1, w_, h_, raw.step, out.data), "vpl_preprocess_batch");
    return out;
  }

  std::vector<float> mapx_, mapy_;
  int w_ = 0, h_ = 0;
  float fx_ = 1, fy_ = 1, cx_ = 0, cy_ = 0;
  bool pre_set_ = false;
  VplContext* pre_ = nullptr;
  int allfeature_cnt;
};

}  // namespace ref
}  // namespace vplines
