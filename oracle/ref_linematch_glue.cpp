/*
 * oracle/ref_linematch_glue.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" door into the reference's own line matcher: LineMatching::Matching
 * (/root/reference/line_matching/src/line_matching.cpp) with its per-point tracker
 * LKTrackerInvoker2D (lk_tracker_invoker_2d.cpp) and getImageNormParams / the KLT constructor
 * (first 40 lines of klt.cpp, cut out by the Makefile into oracle/_ref/), all compiled unmodified
 * against oracle/cvshim.  The rest of klt.cpp (other trackers, SSE Scharr, affine models) is not
 * built; the one member Matching() needs from it, KLT::calc2D, is defined HERE: it builds the two
 * pyramids with the oracle's restatement of cv::buildOpticalFlowPyramid / KLT::calcSharrDeriv
 * (orc_linematch.c, pinned against cv2 4.13) and then runs the reference's LKTrackerInvoker2D
 * over all points level by level, exactly the loop of klt.cpp:598-627.
 */
#include <opencv2/opencv.hpp>

#include <thread>

#define private public /* reads LineMatching::status_/errors_/..., KLT::winSize_/... (access only) */
#include "line_matching.h"
#undef private

extern "C" {
#include "vpl_oracle.h"
}

/* ---- KLT members that live in the unbuilt part of klt.cpp ------------------------------------- */
void KLT::calc(cv::InputArray, cv::InputArray, cv::InputArray, cv::InputOutputArray, cv::OutputArray, cv::OutputArray) {
  CVSHIM_UNSUPPORTED("KLT::calc (not on the Matching path)");
}

void KLT::calc2D(cv::InputArray _prevImg, cv::InputArray _nextImg, cv::InputArray _prevPts,
                 cv::InputOutputArray _nextPts, cv::OutputArray _status, cv::OutputArray _err,
                 bool _illumination_adapt, const std::vector<cv::Mat>* _affines) {
  const cv::Mat& prev = *_prevImg.m;
  const cv::Mat& next = *_nextImg.m;
  std::vector<cv::Point2f>& prevPts = *_prevPts.vp;
  std::vector<cv::Point2f>& nextPts = *_nextPts.vp;
  std::vector<uchar>& status = *_status.vu;
  std::vector<float>& err = *_err.vf;
  const int npoints = (int)prevPts.size();
  if (npoints == 0) { nextPts.clear(); status.clear(); err.clear(); return; }  /* klt.cpp:503-508 */
  if (!(flags_ & cv::OPTFLOW_USE_INITIAL_FLOW)) nextPts.resize(npoints);       /* :510-511 */
  status.resize(npoints);
  for (int i = 0; i < npoints; i++) status[i] = true;                          /* :525-526 */
  err.resize(npoints);                                                         /* :528-533 */
  CV_Assert(prev.isContinuous() && next.isContinuous() && prev.type() == CV_8UC1);
  OrcKltLevel LI[8], LJ[8];
  /* maxLevel_ = buildOpticalFlowPyramid(...) on both images, klt.cpp:589-595 (the member is overwritten) */
  maxLevel_ = orc_klt_build_levels(prev.data, prev.cols, prev.rows, winSize_.width, maxLevel_, 1, LI);
  maxLevel_ = orc_klt_build_levels(next.data, next.cols, next.rows, winSize_.width, maxLevel_, 0, LJ);
  for (int level = maxLevel_; level >= 0; level--) {                           /* :602-627 */
    OrcKltLevel& a = LI[level];
    OrcKltLevel& b = LJ[level];
    cv::Mat I(a.h, a.w, CV_8UC1, a.img + (size_t)a.pad * a.stride + a.pad, (size_t)a.stride);
    cv::Mat derivI(a.h, a.w, CV_16SC2, a.deriv + ((size_t)a.pad * a.stride + a.pad) * 2, (size_t)a.stride * 4);
    cv::Mat J(b.h, b.w, CV_8UC1, b.img + (size_t)b.pad * b.stride + b.pad, (size_t)b.stride);
    /* parallel_for_(Range(0, npoints), LKTrackerInvoker2D(...)): one stripe */
    LKTrackerInvoker2D(I, derivI, J, prevPts.data(), nextPts.data(), status.data(), err.data(), winSize_, criteria_,
                       level, maxLevel_, flags_, (float)minEigThreshold_, _illumination_adapt, _affines)(
        cv::Range(0, npoints));
  }
  orc_klt_free_levels(LI, maxLevel_);
  orc_klt_free_levels(LJ, maxLevel_);
}

extern "C" {

typedef struct {
  float endpoint[4];
  double equation[3];
  float center[2];
  float length;
  float pad_;
} RefLine2;

static void to_lines(const RefLine2* in, int n, std::vector<Line>& out) {
  out.resize(n);
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 4; k++) out[i].line_endpoint[k] = in[i].endpoint[k];
    for (int k = 0; k < 3; k++) out[i].line_equation[k] = in[i].equation[k];
    out[i].center[0] = in[i].center[0];
    out[i].center[1] = in[i].center[1];
    out[i].length = in[i].length;
  }
}

/* LineMatching().Matching(img_ref, img_cur, lines_ref, lines_cur, out, NULL, NULL, NULL, illum, topo, 0)
 * -- the call of feature_tracker/src/line_feature_tracker.cpp:299-310.  Returns Matching()'s bool.
 * Optional per-anchor outputs (cap_kp entries each). */
int ref_line_matching(const uint8_t* img_ref, const uint8_t* img_cur, int w, int h, const RefLine2* lines_ref, int n_ref,
                      const RefLine2* lines_cur, int n_cur, int illumination_adapt, int topological_filter,
                      int32_t* ref_to_cur, float* kps_ref, float* kps_cur, uint8_t* status, float* err,
                      int32_t* kp2line_cur, int cap_kp, int* n_kp) {
  LineMatching lm;
  std::vector<Line> lr, lc;
  to_lines(lines_ref, n_ref, lr);
  to_lines(lines_cur, n_cur, lc);
  cv::Mat a(h, w, CV_8UC1, (void*)img_ref), b(h, w, CV_8UC1, (void*)img_cur);
  std::vector<int> r2c;
  bool ok = lm.Matching(a, b, lr, lc, r2c, *(cv::Mat*)NULL, *(cv::Mat*)NULL, *(cv::Mat*)NULL, illumination_adapt != 0,
                        topological_filter != 0, 0, 0, 0);
  if (n_kp) *n_kp = 0;
  if (!ok) return 0;
  for (int i = 0; i < n_ref; i++) ref_to_cur[i] = r2c[i];
  int n = (int)lm.kps_ref_.size();
  if (n_kp) *n_kp = n;
  for (int i = 0; i < n && i < cap_kp; i++) {
    if (kps_ref) { kps_ref[2 * i] = lm.kps_init_[i].x; kps_ref[2 * i + 1] = lm.kps_init_[i].y; }
    if (kps_cur) { kps_cur[2 * i] = lm.kps_cur_[i].x; kps_cur[2 * i + 1] = lm.kps_cur_[i].y; }
    if (status) status[i] = lm.status_[i];
    if (err) err[i] = lm.errors_[i];
    if (kp2line_cur) kp2line_cur[i] = lm.kp2line_cur_[i];
  }
  return 1;
}

/* The tracker's per-frame hot loop on the host (feature_tracker/src/line_feature_tracker.cpp:87, :115):
 * EDline on every frame + Matching(prev, cur) on every consecutive pair; frames in contiguous chunks
 * over n_threads threads with a one-frame halo, one detector + matcher per thread.  Returns the
 * number of matched lines (timing entry). */
long long ref_linefront_sequence_mt(const uint8_t* frames, int n_frames, int w, int h, const EDLineParam* p, int smoothed,
                                    int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::vector<long long> tot(n_threads, 0);
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) {
    th.emplace_back([&, t]() {
      EDLineDetector det(*p);
      LineMatching lm;
      std::vector<Line> prev, cur;
      std::vector<int> r2c;
      int lo = (int)((long long)n_frames * t / n_threads), hi = (int)((long long)n_frames * (t + 1) / n_threads);
      int start = lo > 0 ? lo - 1 : 0;
      for (int f = start; f < hi; f++) {
        cv::Mat image(h, w, CV_8UC1, (void*)(frames + (size_t)f * w * h));
        det.edges_.xCors.clear(); det.edges_.yCors.clear(); det.edges_.sId.clear(); det.edges_.numOfEdges = 0;
        det.EDline(image, cur, smoothed != 0);
        if (f > start) {
          cv::Mat pimg(h, w, CV_8UC1, (void*)(frames + (size_t)(f - 1) * w * h));
          if (lm.Matching(pimg, image, prev, cur, r2c, *(cv::Mat*)NULL, *(cv::Mat*)NULL, *(cv::Mat*)NULL, true, true, 0, 0, 0))
            for (int v : r2c) tot[t] += v >= 0;
        }
        prev.swap(cur);
      }
    });
  }
  for (auto& x : th) x.join();
  long long s = 0;
  for (auto v : tot) s += v;
  return s;
}

}  // extern "C"
