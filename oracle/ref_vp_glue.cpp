/*
 * oracle/ref_vp_glue.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" door into the reference's own vanishing-point stage, compiled from
 * /root/reference/feature_tracker/src/vanishing_point_detection.cpp (unmodified, where it lies)
 * against oracle/cvshim (OpenCV names, plus stand-ins for the Eigen / glog / camodocal /
 * line_descriptor headers it includes -- each says what it reproduces).  Output:
 * oracle/_ref/libref_vp.so.  Used to pin oracle/orc_vp.c, to generate tests/golden/ref_vp.npz and as
 * the "reference" CPU baseline of bench.py's V1 workload.  Never loaded by the product.
 *
 * The reference seeds rand() with time(NULL) on every call (vanishing_point_detection.cpp:107).  This
 * library defines time() itself (linked with -Bsymbolic-functions, so the reference's call binds
 * here) and returns the seed the caller asked for: the reference's code then runs as written, on a
 * clock we control.  rand()/srand() are the C library's.
 */
#include <ctime>
#include <vector>

#include "vanishing_point_detection.h"

static volatile time_t g_fixed_time = 0; /* read from the std::thread the reference starts (:49): not thread_local */
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
} RefVpLine;

static std::vector<Line> to_lines(const RefVpLine* l, int n) {
  std::vector<Line> v((size_t)n);
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 4; ++k) v[i].line_endpoint[k] = l[i].endpoint[k];
    for (int k = 0; k < 3; ++k) v[i].line_equation[k] = l[i].equation[k];
    v[i].center[0] = l[i].center[0]; v[i].center[1] = l[i].center[1];
    v[i].length = l[i].length;
  }
  return v;
}

/* run_vanishing_point_detection on a fresh object after init(f, cx, cy, 0.5).  frame_count > 0: the
 * object has made one call before (the member only distinguishes 0 from the rest, :337-349).
 * rand() is process-global: callers must not run this concurrently from several threads. */
int ref_vp_detect(const RefVpLine* lines, int n_lines, const RefVpLine* all_lines, int n_all, float f, float cx,
                  float cy, unsigned seed, int frame_count, double* vps, int32_t* vp_idx) {
  if (n_lines < 3) return -1;
  vanishing_point_detection d;
  d.init(f, cx, cy, 0.5);
  std::vector<Line> L = to_lines(lines, n_lines), A = to_lines(all_lines, n_all);
  cv::Mat img;
  g_fixed_time = (time_t)seed;
  if (frame_count > 0) {
    std::vector<Eigen::Vector3d> v0;
    std::vector<int> id0;
    std::vector<Line> none; /* no lines to classify: the warm-up call cannot reach the lx[idx] reads */
    d.run_vanishing_point_detection(img, L, none, v0, id0);
  }
  std::vector<Eigen::Vector3d> v;
  std::vector<int> ids;
  d.run_vanishing_point_detection(img, L, A, v, ids);
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) vps[3 * i + k] = v[i](k);
  for (int i = 0; i < n_all; ++i) vp_idx[i] = ids[(size_t)i];
  return 0;
}

/* n_frames frames one after another on one object (timing + sequence parity): frame i uses
 * lines + i*cap (counts[i] of them) as both `lines` and `all_lines`, seed seeds[i]; frame_count0 > 0: the
 * object has made a call before frame 0. */
long long ref_vp_sequence(const RefVpLine* lines, const int32_t* counts, int n_frames, int cap, float f, float cx,
                          float cy, const uint32_t* seeds, int frame_count0, double* vps, int32_t* vp_idx) {
  vanishing_point_detection d;
  d.init(f, cx, cy, 0.5);
  cv::Mat img;
  long long labelled = 0;
  if (frame_count0 > 0) { /* an object that has made a call before: warm-up call with nothing to classify */
    for (int i = 0; i < n_frames; ++i) {
      if (counts[i] < 3) continue;
      std::vector<Line> L = to_lines(lines + (size_t)i * cap, counts[i]), none;
      std::vector<Eigen::Vector3d> v0;
      std::vector<int> id0;
      g_fixed_time = (time_t)seeds[i];
      d.run_vanishing_point_detection(img, L, none, v0, id0);
      break;
    }
  }
  for (int i = 0; i < n_frames; ++i) {
    if (counts[i] < 3) continue;
    std::vector<Line> L = to_lines(lines + (size_t)i * cap, counts[i]);
    g_fixed_time = (time_t)seeds[i];
    std::vector<Eigen::Vector3d> v;
    std::vector<int> ids;
    d.run_vanishing_point_detection(img, L, L, v, ids);
    for (int a = 0; a < 3; ++a)
      for (int k = 0; k < 3; ++k) vps[9 * (size_t)i + 3 * a + k] = v[a](k);
    for (int k = 0; k < counts[i]; ++k) { vp_idx[(size_t)i * cap + k] = ids[(size_t)k]; labelled += ids[(size_t)k] != 3; }
  }
  return labelled;
}

}  // extern "C"
