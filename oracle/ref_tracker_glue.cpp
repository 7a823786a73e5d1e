/*
 * oracle/ref_tracker_glue.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" door into the reference's own LineFeatureTracker::readImage, compiled from
 * /root/reference/feature_tracker/src/line_feature_tracker.cpp (unmodified, where it lies) together with the
 * reference's line primitives (edline_detector.cpp, line_matching.cpp, lk_tracker_invoker_2d.cpp) and its
 * vanishing-point stage (vanishing_point_detection.cpp) against oracle/cvshim.  Output:
 * oracle/_ref/libref_tracker.so.  Used to pin oracle/orc_tracker.py and to generate tests/golden/ref_tracker.npz.
 * Never loaded by the product.
 *
 * What is NOT the reference's code here: the OpenCV / Eigen / ROS / camodocal stand-ins under oracle/cvshim (cv::remap
 * and cv::CLAHE are the oracle's restatements, pinned against cv2 4.13), KLT::calc2D (ref_linematch_glue.cpp) and
 * time(): the vanishing-point stage seeds rand() with time(NULL) on every call; this library defines time() itself
 * (linked with -Bsymbolic-functions) and returns the seed the caller set for the frame.
 */
#include <ctime>
#include <vector>

#include "linefeature_tracker.h"

/* the globals of feature_tracker/src/parameters.cpp that readImage reads */
int EQUALIZE = 1;
int max_h_lines = 25;
int max_v_lines = 25;
float MIN_LINE_LENGTH = 35.f;
float line_fit_err = 1.8f;

camodocal::CvshimIntrinsics camodocal::g_cvshim_intrinsics = {nullptr, nullptr, 0, 0, 0, 0, 0, 0};

static volatile time_t g_fixed_time = 0;
extern "C" time_t time(time_t* t) {
  if (t) *t = g_fixed_time;
  return g_fixed_time;
}

extern "C" {

typedef struct {
  float endpoint[4];
  double equation[3];
  float center[2];
  float length;
  float pad_;
} RefTrLine;

/* a tracker set up as main() of line_feature_tracker_node.cpp sets it up (:196-207): EDLineParam{5, 1, 30, 5, 2,
 * MIN_LINE_LENGTH, line_fit_err}, LineMatching() defaults, readIntrinsicParameter (maps + K through the camera
 * stand-in), the YAML's EQUALIZE / max_h_lines / max_v_lines. */
void* ref_tracker_create(const float* mapx, const float* mapy, int w, int h, float fx, float fy, float cx, float cy,
                         int equalize, int max_h, int max_v, float min_line_length, float fit_err) {
  EQUALIZE = equalize; max_h_lines = max_h; max_v_lines = max_v;
  MIN_LINE_LENGTH = min_line_length; line_fit_err = fit_err;
  camodocal::g_cvshim_intrinsics = {mapx, mapy, w, h, fx, fy, cx, cy};
  LineFeatureTracker* t = new LineFeatureTracker();
  EDLineParam param = {5, 1.0, 30, 5, 2, MIN_LINE_LENGTH, line_fit_err};
  t->line_detctor = EDLineDetector(param);
  t->line_matching = LineMatching();
  t->readIntrinsicParameter("unused.yaml");
  return t;
}
void ref_tracker_destroy(void* h) { delete (LineFeatureTracker*)h; }

/* readImage(img) with time(NULL) == seed during the call.  Afterwards: *lines_exit, and curframe_'s vecLine /
 * lineID / vps (4 doubles per line, n_vps of them: 0 after the first image) / t_cnt (n_tcnt entries: the raw
 * detection count of the frame, the reference never swaps it with the selection).  Returns the line count, or
 * -(needed) if cap is too small. */
int ref_tracker_read(void* h, const uint8_t* img, int w, int hgt, unsigned seed, int cap, RefTrLine* lines, int32_t* ids,
                     double* vps, int32_t* n_vps, int32_t* t_cnt, int32_t* n_tcnt, int32_t* lines_exit) {
  LineFeatureTracker* t = (LineFeatureTracker*)h;
  cv::Mat m(hgt, w, CV_8UC1, (void*)img);
  g_fixed_time = (time_t)seed;
  t->readImage(m);
  *lines_exit = t->lines_exit ? 1 : 0;
  const FrameLines& F = *t->curframe_;
  const int n = (int)F.vecLine.size();
  if (n > cap || (int)F.t_cnt.size() > cap) return -(n > (int)F.t_cnt.size() ? n : (int)F.t_cnt.size());
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 4; ++k) lines[i].endpoint[k] = F.vecLine[i].line_endpoint[k];
    for (int k = 0; k < 3; ++k) lines[i].equation[k] = F.vecLine[i].line_equation[k];
    lines[i].center[0] = F.vecLine[i].center[0]; lines[i].center[1] = F.vecLine[i].center[1];
    lines[i].length = F.vecLine[i].length; lines[i].pad_ = 0;
    ids[i] = F.lineID[i];
  }
  *n_vps = (int)F.vps.size();
  for (int i = 0; i < (int)F.vps.size(); ++i)
    for (int k = 0; k < 4; ++k) vps[4 * i + k] = F.vps[(size_t)i](k);
  *n_tcnt = (int)F.t_cnt.size();
  for (int i = 0; i < (int)F.t_cnt.size(); ++i) t_cnt[i] = F.t_cnt[(size_t)i];
  return n;
}

}  /* extern "C" */
