// line_matching_b200.hpp -- the reference's OWN line primitives interface, over libvplines_b200.so.
//
// Header-only mirror of the classes LineFeatureTracker holds
// (/root/reference/feature_tracker/include/linefeature_tracker.h:81-83):
//     struct Line                                   line_matching/src/line.h:8-17
//     struct EDLineParam                            line_matching/src/edline_detector.h:32-40
//     class  EDLineDetector  { int  EDline(cv::Mat&, std::vector<Line>&, bool smoothed); }
//                                                   edline_detector.h:79-81, edline_detector.cpp:1176
//     class  LineMatching    { bool Matching(img_ref, img_cur, lines_ref, lines_cur, line_ref_to_line_cur,
//                                            K_ref, K_cur, T_cur_ref, illumination_adapt,
//                                            topological_filter, debug_show, ...); }
//                                                   line_matching.h:21-33, line_matching.cpp:605
//     class  vanishing_point_detection { void init(f, cx, cy, noiseRatio);
//                                        void run_vanishing_point_detection(img, lines, all_lines, vps, local_vp_ids); }
//                                                   feature_tracker/include/vanishing_point_detection.h:44-84
// Same names, argument meaning and return values, so the tracker's two seams
// (edline_detect / match_line_match, line_feature_tracker.cpp:291-321) compile against it
// unchanged -- put `using namespace vplines::ref;` (or the two `using` lines below) where the
// reference includes "line_matching.h".  Results are bit-identical to the reference's CPU code run
// on one thread (tests/test_gpu_edlines.py, tests/test_gpu_linematch.py); with several OpenCV
// threads the reference's own line ORDER is a race (edline_detector.cpp:1081-1083), here it is
// always (edge chain, position).  There is no CPU fallback: every call throws
// std::runtime_error(vpl_last_error()) when the device path fails.
//
// vanishing_point_detection: the reference seeds rand() with time(NULL) on every call
// (vanishing_point_detection.cpp:107); so does this class, through seed_source (default: time(NULL)),
// which a caller can replace to make runs repeatable.  Given the seed, the result is the reference's,
// computed with correctly rounded atan/acos/sin/cos instead of libm's (tests/test_gpu_vp.py: labels
// identical, vanishing points within 1e-15); where the reference would read its lx[] out of range
// (:438-441, :456-459) no query is made (last_status() == 1).
//
// Deliberate differences, all outside what the tracker uses:
//   - K_ref / K_cur / T_cur_ref are accepted and ignored: Matching() itself never reads them for the
//     2-D tracker (the tracker passes null references, line_feature_tracker.cpp:304-306);
//   - debug_show > 0 (OpenCV windows) is not available;
//   - EDline on a frame where EdgeDrawing fails returns 1 with no lines, where the reference goes on
//     with the previous frame's stale edge chains (edline_detector.cpp:1180: `!(-1)` is false).
#pragma once
#include <algorithm>
#include <array>
#include <cstring>
#include <ctime>
#include <stdexcept>
#include <string>
#include <vector>

#include "line_descriptor.hpp"  // the minimal cv::Mat (or the real one with -DVPL_WITH_OPENCV) + vpl_capi.h

namespace vplines {
namespace ref {

struct Line {  // line.h:8-17 (the kps / kps_init / dirs members are debug-only and not carried)
  std::array<float, 4> line_endpoint;
  std::array<double, 3> line_equation;
  std::array<float, 2> center;
  float length;
};

struct EDLineParam {  // edline_detector.h:32-40
  int ksize;
  float sigma;
  float gradientThreshold;
  float anchorThreshold;
  int scanIntervals;
  int minLineLen;
  double lineFitErrThreshold;
};

namespace detail {
// one context per thread for the reference-named classes (EDLines + KLT matching only: lsd_path = 0)
struct RefCtx {
  VplContext* h = nullptr;
  int w = 0, hh = 0;
  VplEDLineParam edp{};
  VplLineMatchParam lmp{};
  bool ed_set = false, lm_set = false;
  const void* vp_owner = nullptr;  // the vanishing_point_detection object whose intrinsics are configured
  ~RefCtx() { if (h) vpl_destroy(h); }
  static void check(VplContext* c, int r, const char* what) {
    if (r != VPL_OK) throw std::runtime_error(std::string(what) + ": " + vpl_last_error(c));
  }
  VplContext* get(int width, int height) {
    if (h && width <= w && height <= hh) return h;
    VplConfig c;
    vpl_default_config(&c);
    c.max_width = std::max(width, w);
    c.max_height = std::max(height, hh);
    c.max_lines = 2048;
    c.max_batch = 2;
    c.num_slots = 1;
    c.lsd_path = 0;
    VplContext* n = nullptr;
    int r = vpl_create(&c, &n);
    if (r != VPL_OK) throw std::runtime_error(std::string("vpl_create: ") + vpl_last_error(nullptr));
    if (h) vpl_destroy(h);
    h = n; w = c.max_width; hh = c.max_height;
    ed_set = lm_set = false;
    vp_owner = nullptr;
    return h;
  }
  static bool same(const VplEDLineParam& a, const VplEDLineParam& b) {
    return a.ksize == b.ksize && a.sigma == b.sigma && a.gradientThreshold == b.gradientThreshold &&
           a.anchorThreshold == b.anchorThreshold && a.scanIntervals == b.scanIntervals && a.minLineLen == b.minLineLen &&
           a.lineFitErrThreshold == b.lineFitErrThreshold;
  }
};
inline RefCtx& rctx() {
  thread_local RefCtx c;
  return c;
}
inline void need_u8(const cv::Mat& m) {
  if (m.depth() != 0 || m.channels() != 1 || m.empty()) throw std::runtime_error("line_matching_b200: CV_8UC1 image expected");
}
}  // namespace detail

class EDLineDetector {
 public:
  EDLineDetector() {  // edline_detector.cpp:18-29
    p_.ksize = 5; p_.sigma = 1.0f; p_.gradientThreshold = 80; p_.anchorThreshold = 2; p_.scanIntervals = 2;
    p_.minLineLen = 15; p_.lineFitErrThreshold = 1.4;
  }
  explicit EDLineDetector(EDLineParam param) {  // edline_detector.cpp:30-40
    p_.ksize = param.ksize; p_.sigma = param.sigma; p_.gradientThreshold = param.gradientThreshold;
    p_.anchorThreshold = param.anchorThreshold; p_.scanIntervals = param.scanIntervals; p_.minLineLen = param.minLineLen;
    p_.lineFitErrThreshold = param.lineFitErrThreshold;
  }
  // 1 = done (also when the frame has no lines), as the reference; throws on a device failure
  int EDline(cv::Mat& image, std::vector<Line>& lines, bool smoothed = false) {
    detail::need_u8(image);
    detail::RefCtx& c = detail::rctx();
    VplContext* h = c.get(image.cols, image.rows);
    if (!c.ed_set || !detail::RefCtx::same(c.edp, p_)) {
      detail::RefCtx::check(h, vpl_edlines_configure(h, &p_), "vpl_edlines_configure");
      c.edp = p_; c.ed_set = true;
    }
    const uint8_t* ptr = image.data;
    buf_.resize(2048);
    int32_t count = 0, status = 0;
    detail::RefCtx::check(h, vpl_edlines_detect_batch(h, &ptr, 1, image.cols, image.rows, image.step, smoothed ? 1 : 0,
                                                      buf_.data(), &count, (int)buf_.size(), &status),
                          "vpl_edlines_detect_batch");
    last_status_ = status;
    lines.clear();
    lines.reserve((size_t)count);
    for (int i = 0; i < count; ++i) {
      Line l;
      for (int k = 0; k < 4; ++k) l.line_endpoint[(size_t)k] = buf_[(size_t)i].endpoint[k];
      for (int k = 0; k < 3; ++k) l.line_equation[(size_t)k] = buf_[(size_t)i].equation[k];
      l.center[0] = buf_[(size_t)i].center[0]; l.center[1] = buf_[(size_t)i].center[1];
      l.length = buf_[(size_t)i].length;
      lines.push_back(l);
    }
    return 1;
  }
  int lastEdgeDrawingStatus() const { return last_status_; }  // EdgeDrawing's 1 / -1 for the last frame

 private:
  VplEDLineParam p_;
  std::vector<VplLine> buf_;
  int last_status_ = 1;
};

class LineMatching {
 public:
  LineMatching(int step = 10, float closest_line_threshold = 0.5f, float line_matching_ratio = 0.4f,
               float line_distance_error_ratio = 3, float klt_error_threshold = 40) {  // line_matching.cpp:3-15
    vpl_linematch_default_param(&p_);
    p_.step = step; p_.closest_line_threshold = closest_line_threshold; p_.line_matching_ratio = line_matching_ratio;
    p_.line_distance_error_ratio = line_distance_error_ratio; p_.klt_error_threshold = klt_error_threshold;
  }
  // false when a line set is empty (line_matching.cpp:621), otherwise true with
  // line_ref_to_line_cur[i] = index of the current line matched to reference line i, or -1
  bool Matching(const cv::Mat& img_ref, const cv::Mat& img_cur, const std::vector<Line>& lines_ref,
                const std::vector<Line>& lines_cur, std::vector<int>& line_ref_to_line_cur,
                const cv::Mat& /*K_ref*/ = *(const cv::Mat*)nullptr, const cv::Mat& /*K_cur*/ = *(const cv::Mat*)nullptr,
                const cv::Mat& /*T_cur_ref*/ = *(const cv::Mat*)nullptr, const bool illumination_adapt = false,
                const bool topological_filter = true, const int debug_show = 0, const int = 0, const int = 0) {
    if (lines_ref.empty() || lines_cur.empty()) return false;
    if (debug_show > 0) throw std::runtime_error("LineMatching::Matching: debug_show is not available on the device path");
    detail::need_u8(img_ref);
    detail::need_u8(img_cur);
    if (img_ref.cols != img_cur.cols || img_ref.rows != img_cur.rows || img_ref.step != img_cur.step)
      throw std::runtime_error("LineMatching::Matching: the two images must have the same geometry");
    detail::RefCtx& c = detail::rctx();
    VplContext* h = c.get(img_ref.cols, img_ref.rows);
    VplLineMatchParam p = p_;
    p.illumination_adapt = illumination_adapt ? 1 : 0;
    p.topological_filter = topological_filter ? 1 : 0;
    if (!c.lm_set || std::memcmp(&c.lmp, &p, sizeof(p)) != 0) {
      detail::RefCtx::check(h, vpl_linematch_configure(h, &p), "vpl_linematch_configure");
      c.lmp = p; c.lm_set = true;
    }
    const int cap = (int)std::max(lines_ref.size(), lines_cur.size());
    std::vector<VplLine> a((size_t)cap), b((size_t)cap);
    pack(lines_ref, a);
    pack(lines_cur, b);
    int32_t na = (int32_t)lines_ref.size(), nb = (int32_t)lines_cur.size();
    std::vector<int32_t> out((size_t)cap, -1);
    const uint8_t *pr = img_ref.data, *pc = img_cur.data;
    detail::RefCtx::check(h, vpl_linematch_batch(h, &pr, &pc, 1, img_ref.cols, img_ref.rows, img_ref.step, a.data(), &na,
                                                 b.data(), &nb, cap, out.data()),
                          "vpl_linematch_batch");
    line_ref_to_line_cur.assign(out.begin(), out.begin() + na);
    return true;
  }

 private:
  static void pack(const std::vector<Line>& in, std::vector<VplLine>& out) {
    for (size_t i = 0; i < in.size(); ++i) {
      for (int k = 0; k < 4; ++k) out[i].endpoint[k] = in[i].line_endpoint[(size_t)k];
      for (int k = 0; k < 3; ++k) out[i].equation[k] = in[i].line_equation[(size_t)k];
      out[i].center[0] = in[i].center[0]; out[i].center[1] = in[i].center[1];
      out[i].length = in[i].length;
      out[i].reserved = 0;
    }
  }
  VplLineMatchParam p_;
};

struct Vector3d {  // stands in for Eigen::Vector3d where Eigen is not available (any type with this
  double v[3];     // constructor and operator() works: the method below is a template on it)
  Vector3d() : v{0, 0, 0} {}
  Vector3d(double a, double b, double c) : v{a, b, c} {}
  double& operator()(int i) { return v[i]; }
  const double& operator()(int i) const { return v[i]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
};

class vanishing_point_detection {  // feature_tracker/include/vanishing_point_detection.h:44-84
 public:
  vanishing_point_detection() {}
  void init(float _f, float _cx, float _cy, double _noiseRatio) {  // vanishing_point_detection.cpp:29-34
    f = _f; pp_x = _cx; pp_y = _cy; noiseRatio = _noiseRatio;
    f_ = _f; cx_ = _cx; cy_ = _cy;
    configured_ = false;
  }
  // vps: three unit vectors; local_vp_ids: one label per entry of all_lines appended (0..2, 3 = none), as the
  // reference's push_back.  Throws on a device failure and where the reference itself cannot proceed
  // (fewer than 2 lines; no pair of lines that intersect).
  template <class Vec3>
  void run_vanishing_point_detection(const cv::Mat& /*img: only drawn on*/, std::vector<Line>& lines,
                                     std::vector<Line>& all_lines, std::vector<Vec3>& vps, std::vector<int>& local_vp_ids) {
    detail::RefCtx& c = detail::rctx();
    VplContext* h = c.get(std::max(c.w, 64), std::max(c.hh, 64));
    if (!configured_ || c.vp_owner != this) {
      detail::RefCtx::check(h, vpl_vp_configure(h, f_, cx_, cy_), "vpl_vp_configure");
      configured_ = true; c.vp_owner = this;
    }
    if (lines.size() > 2048 || all_lines.size() > 2048) throw std::runtime_error("vanishing_point_detection: more than 2048 lines");
    const int cap = (int)std::max<size_t>(1, std::max(lines.size(), all_lines.size()));
    std::vector<VplLine> a((size_t)cap), b((size_t)cap);
    pack(lines, a);
    pack(all_lines, b);
    const int32_t na = (int32_t)lines.size(), nb = (int32_t)all_lines.size();
    const uint32_t seed = seed_source ? seed_source() : (uint32_t)std::time(nullptr);
    double out[9];
    std::vector<int32_t> idx((size_t)cap, 3);
    int32_t status = 0;
    detail::RefCtx::check(h, vpl_vp_detect_batch(h, a.data(), &na, b.data(), &nb, 1, cap, &seed, frame_count, out, idx.data(),
                                                 nullptr, &status),
                          "vpl_vp_detect_batch");
    last_status_ = status;
    if (status < 0)
      throw std::runtime_error(status == -1 ? "vanishing_point_detection: fewer than 2 lines"
                                            : "vanishing_point_detection: no two lines intersect");
    vps.clear();
    for (int k = 0; k < 3; ++k) vps.push_back(Vec3(out[3 * k], out[3 * k + 1], out[3 * k + 2]));
    for (int32_t i = 0; i < nb; ++i) local_vp_ids.push_back(idx[(size_t)i]);
    frame_count++;
  }
  int last_status() const { return last_status_; }
  uint32_t (*seed_source)() = nullptr;  // nullptr: time(NULL), as the reference

  double pp_x = 0, pp_y = 0;  // the reference's cv::Point2d pp
  double f = 0;
  double noiseRatio = 0;

 private:
  static void pack(const std::vector<Line>& in, std::vector<VplLine>& out) {
    for (size_t i = 0; i < in.size(); ++i) {
      for (int k = 0; k < 4; ++k) out[i].endpoint[k] = in[i].line_endpoint[(size_t)k];
      for (int k = 0; k < 3; ++k) out[i].equation[k] = in[i].line_equation[(size_t)k];
      out[i].center[0] = in[i].center[0]; out[i].center[1] = in[i].center[1];
      out[i].length = in[i].length;
      out[i].reserved = 0;
    }
  }
  float f_ = 0, cx_ = 0, cy_ = 0;
  bool configured_ = false;
  int frame_count = 0;
  int last_status_ = 0;
};

}  // namespace ref
}  // namespace vplines
