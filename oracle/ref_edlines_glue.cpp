/*
 * oracle/ref_edlines_glue.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" door into the reference's own EDLines detector, compiled from
 * /root/reference/line_matching/src/edline_detector.cpp (unmodified, where it lies) against
 * oracle/cvshim (see that header for what is substituted).  Output: part of oracle/_ref/libref_linefront.so.
 * Used to (1) pin oracle/orc_edlines.c, (2) generate tests/golden/ref_edlines.npz
 * (tests/golden/make_golden_edlines.py) and (3) serve as the "reference" CPU baseline of
 * bench.py's EDLines workload.  Never loaded by the product.
 */
#include <opencv2/opencv.hpp>

#include <thread>

#define private public /* the glue reads EDLineDetector::edges_ (same layout, access only) */
#include "edline_detector.h"
#undef private

extern "C" {

/* same layout as OrcLine / VplLine: Line's numeric fields (line.h:8-12) */
typedef struct {
  float endpoint[4];
  double equation[3];
  float center[2];
  float length;
} RefLine;

typedef struct {
  int ksize;
  float sigma;
  float gradientThreshold;
  float anchorThreshold;
  int scanIntervals;
  int minLineLen;
  double lineFitErrThreshold;
} RefEDLineParam;

static EDLineParam to_param(const RefEDLineParam* p) {
  EDLineParam q = {p->ksize, p->sigma, p->gradientThreshold, p->anchorThreshold, p->scanIntervals,
                   p->minLineLen, p->lineFitErrThreshold};
  return q;
}

static void clear_edges(EDLineDetector& d) {
  /* EDline() ignores EdgeDrawing()'s -1 (edline_detector.cpp:1180: `!(-1)` is false) and then
   * walks whatever edges_ holds; a fresh or failed frame must therefore see an empty set. */
  d.edges_.xCors.clear();
  d.edges_.yCors.clear();
  d.edges_.sId.clear();
  d.edges_.numOfEdges = 0;
}

/* One frame through EDLineDetector::EDline (edline_detector.cpp:1176).  Returns the number of
 * lines (all of them are counted, at most cap are written).  Optional stage outputs:
 * chain_xy (x | y << 16 per edge pixel, cap_px entries), chain_sid (cap_chains + 1), n_px,
 * n_chains. */
int ref_edline_detect(const uint8_t* img, int w, int h, const RefEDLineParam* p, int smoothed,
                      RefLine* out, int cap, uint32_t* chain_xy, int cap_px, uint32_t* chain_sid,
                      int cap_chains, int* n_px, int* n_chains) {
  EDLineDetector det(to_param(p));
  clear_edges(det);
  cv::Mat image(h, w, CV_8UC1, (void*)img);
  std::vector<Line> lines;
  det.EDline(image, lines, smoothed != 0);
  int n = (int)lines.size();
  for (int i = 0; i < n && i < cap; i++) {
    for (int k = 0; k < 4; k++) out[i].endpoint[k] = lines[i].line_endpoint[k];
    for (int k = 0; k < 3; k++) out[i].equation[k] = lines[i].line_equation[k];
    out[i].center[0] = lines[i].center[0];
    out[i].center[1] = lines[i].center[1];
    out[i].length = lines[i].length;
  }
  int ne = (int)det.edges_.numOfEdges;
  int np = ne ? (int)det.edges_.sId[ne] : 0;
  if (n_chains) *n_chains = ne;
  if (n_px) *n_px = np;
  if (chain_xy)
    for (int i = 0; i < np && i < cap_px; i++) chain_xy[i] = det.edges_.xCors[i] | (det.edges_.yCors[i] << 16);
  if (chain_sid)
    for (int i = 0; i <= ne && i <= cap_chains; i++) chain_sid[i] = ne ? det.edges_.sId[i] : 0;
  return n;
}

/* Timing entry: n_frames frames (contiguous, w*h each) split over n_threads host threads, one
 * EDLineDetector per thread reused across its frames (as the tracker node reuses its detector,
 * feature_tracker/src/line_feature_tracker_node.cpp:203).  Returns the total number of lines. */
long long ref_edline_sequence_mt(const uint8_t* frames, int n_frames, int w, int h,
                                 const RefEDLineParam* p, int smoothed, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::vector<long long> tot(n_threads, 0);
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) {
    th.emplace_back([&, t]() {
      EDLineDetector det(to_param(p));
      clear_edges(det);
      std::vector<Line> lines;
      int lo = (int)((long long)n_frames * t / n_threads), hi = (int)((long long)n_frames * (t + 1) / n_threads);
      for (int f = lo; f < hi; f++) {
        cv::Mat image(h, w, CV_8UC1, (void*)(frames + (size_t)f * w * h));
        clear_edges(det);
        det.EDline(image, lines, smoothed != 0);
        tot[t] += (long long)lines.size();
      }
    });
  }
  for (auto& x : th) x.join();
  long long s = 0;
  for (auto v : tot) s += v;
  return s;
}

}  // extern "C"
