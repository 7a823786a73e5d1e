/*
 * vpl_oracle.h -- CPU ORACLE for the line front end (LSD -> LBD -> Hamming kNN).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  The product path
 * (vplines-slam_b200/csrc) never links it and has no CPU fallback.
 *
 * What it restates (the north-star path is NOT in /root/reference, see SURVEY.md
 * section 0; the arithmetic lives in un-vendored third-party code):
 *   - opencv imgproc  (pinned by the reference: "OpenCV 3.4.2", README.md:5;
 *     find_package(OpenCV 3.4 REQUIRED), vins_estimator/CMakeLists.txt:18):
 *     GaussianBlur / pyrDown / resize(INTER_LINEAR_EXACT) / Sobel / LSD (lsd.cpp).
 *     Only cv2 4.13 is obtainable in this image, so cv2 4.13 semantics are the
 *     pin (its rect_nfa differs from 3.4.2's): every function below is checked
 *     bit-for-bit against cv2 4.13 by tests/test_oracle_*.py and by the golden
 *     vectors under tests/golden/ (made by tests/golden/make_golden.py).
 *   - opencv_contrib 3.4 line_descriptor (LSDDetector.cpp, binary_descriptor.cpp,
 *     binary_descriptor_matcher.cpp): restated from the published algorithm.
 *     cv2.line_descriptor is absent here => LBD / KeyLine packing are
 *     "PARITY UNPINNED" (no upstream vectors exist; the reference has no test
 *     on this path).  Hamming kNN is pinned against
 *     cv2.BFMatcher(NORM_HAMMING).
 *   - the reference's own nfa()/log_gamma (line_matching/src/edline_detector.h:
 *     210-348) are the same formulas LSD uses; cited where followed.
 */
#ifndef VPL_ORACLE_H
#define VPL_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* KeyLine exactly as opencv_contrib line_descriptor lays it out (17 x 4 B). */
typedef struct {
  float angle;
  int32_t class_id;
  int32_t octave;
  float pt_x, pt_y;
  float response;
  float size;
  float startPointX, startPointY, endPointX, endPointY;
  float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
  float lineLength;
  int32_t numOfPixels;
} OrcKeyLine;

/* ---- image primitives (bit-exact vs cv2 4.13) ---------- */
void orc_gaussian_blur5(const uint8_t* src, int w, int h, uint8_t* dst);     /* 5x5 sigma=1   */
void orc_gaussian_blur7_s075(const uint8_t* src, int w, int h, uint8_t* dst);/* 7x7 sigma=.75 */
void orc_resize_08(const uint8_t* src, int w, int h, uint8_t* dst, int* dw, int* dh);
void orc_pyrdown_half(const uint8_t* src, int w, int h, uint8_t* dst);       /* -> (w/2,h/2)  */
void orc_sobel3(const uint8_t* src, int w, int h, int16_t* dx, int16_t* dy);
float orc_fast_atan2(float y, float x);

/* ---- pre-processing in front of the path (readImage, line_feature_tracker.cpp:62-68) ---- */
/* cv::remap(src, dst, mapx, mapy, INTER_LINEAR, BORDER_CONSTANT 0) with CV_32FC1 maps of dw x dh */
void orc_remap_linear(const uint8_t* src, int w, int h, const float* mapx, const float* mapy, int dw,
                      int dh, uint8_t* dst);
void orc_remap_weight_table(uint16_t* tab /* 1024 x 4 */);
/* cv::createCLAHE(clip_limit, Size(tiles,tiles))->apply on CV_8UC1 */
void orc_clahe(const uint8_t* src, int w, int h, double clip_limit, int tiles, uint8_t* dst);

/* ---- LSD (cv::LineSegmentDetector, LSD_REFINE_ADV by default) ------------- */
/* refine: 0 NONE, 1 STD, 2 ADV.  scale08: 1 => internal 0.8 scaling (default),
 * 0 => scale 1.  Outputs up to cap segments; returns the number found (may
 * exceed cap, in which case only cap were written). */
int orc_lsd_detect(const uint8_t* img, int w, int h, int refine, int scale08,
                   float* seg4, double* width, double* prec, double* nfa, int cap);

/* Candidate rectangles after refine(), before rect_improve (16 doubles each: x1 y1 x2 y2
 * width x y theta dx dy prec p | final nfa, accepted, region size, seed pixel index). */
int orc_lsd_candidates(const uint8_t* img, int w, int h, double* out, int cap);

/* Stage dump (0.8-scaled image, angle in degrees / -1024, ordered defined pixels);
 * returns the number of ordered pixels.  Any output may be NULL. */
int orc_lsd_stages(const uint8_t* img, int w, int h, uint8_t* scaled, float* ang_deg, int* order,
                   int* ws, int* hs);

/* ---- LSDDetector::detect (opencv_contrib LSDDetector.cpp) ----------------- */
int orc_lsd_detector_detect(const uint8_t* img, int w, int h, int scale, int num_octaves,
                            int blur_first, OrcKeyLine* out, int cap);

/* ---- BinaryDescriptor::compute (opencv_contrib binary_descriptor.cpp) ----- */
/* desc: n x 32 bytes; fdesc (optional, may be NULL): n x 72 floats. */
int orc_lbd_compute(const uint8_t* img, int w, int h, const OrcKeyLine* kl, int n,
                    uint8_t* desc, float* fdesc);

/* ---- BinaryDescriptorMatcher brute force -------------- */
/* idx/dist: nq x k, ascending distance, lowest train index on ties; -1 pad. */
void orc_hamming_knn(const uint8_t* q, int nq, const uint8_t* t, int nt, int k,
                     int32_t* idx, int32_t* dist);

/* ---- the reference's real detector: EDLineDetector::EDline (orc_edlines.c) ---- */
typedef struct { /* EDLineParam, line_matching/src/edline_detector.h:32-40 */
  int ksize;
  float sigma;
  float gradientThreshold;
  float anchorThreshold;
  int scanIntervals;
  int minLineLen;
  double lineFitErrThreshold;
} OrcEDLineParam;
typedef struct { /* the numeric fields of struct Line, line_matching/src/line.h:8-12 (56 bytes) */
  float endpoint[4];
  double equation[3];
  float center[2];
  float length;
  float pad_;
} OrcLine;
/* Returns the number of lines found (at most cap written), in (edge chain, position) order.
 * Optional stage outputs: chain_xy (x | y << 16 per edge pixel; room for 2*(w*h/5)), chain_sid
 * (room for w*h/100 + 1), their counts. */
int orc_edline_detect(const uint8_t* img, int w, int h, const OrcEDLineParam* p, int smoothed,
                      OrcLine* out, int cap, uint32_t* chain_xy, uint32_t* chain_sid, int* n_px,
                      int* n_chains);
int orc_edge_drawing(const uint8_t* img, int w, int h, const OrcEDLineParam* p, int smoothed,
                     int16_t* dx, int16_t* dy, uint8_t* dir, uint32_t* chain_xy, uint32_t* chain_sid,
                     int* n_px, int* n_chains);
void orc_ed_gradient_maps(const uint8_t* img, int w, int h, int smoothed, int grad_thresh,
                          int16_t* dx, int16_t* dy, int16_t* g, uint8_t* dir);
double orc_ed_nfa(int n, int k, double p, double logNT);
/* n_frames frames over n_threads host threads (timing); returns the total number of lines */
int64_t orc_edline_sequence_mt(const uint8_t* frames, int n_frames, int w, int h,
                               const OrcEDLineParam* p, int smoothed, int n_threads);

/* ---- the reference's real matcher: LineMatching::Matching (orc_linematch.c) ---- */
typedef struct {
  /* LineMatching ctor defaults, line_matching/src/line_matching.h:14-18 */
  int step;
  float closest_line_threshold, line_matching_ratio, line_distance_error_ratio, klt_error_threshold;
  /* KLT as Matching configures it, line_matching.cpp:14, :630-631 */
  int win, max_level, max_count;
  double epsilon;
  float min_eig;
  /* TopologicalFilter defaults, line_matching.h:45-47 */
  float topo_distance_threshold, topo_length_ratio, topo_violation_ratio;
} OrcLineMatchParam;
void orc_lm_default_param(OrcLineMatchParam* p);
/* One level of cv::buildOpticalFlowPyramid(img, winSize, maxLevel, withDerivatives=false) plus the
 * zero-padded Scharr derivative KLT::calc2D adds for the previous image (klt.cpp:598-613). */
typedef struct {
  int w, h, pad;   /* level size and border (= winSize) */
  int stride;      /* w + 2 pad, in pixels */
  uint8_t* img;    /* (h + 2 pad) x stride, BORDER_REFLECT_101 */
  int16_t* deriv;  /* (h + 2 pad) x stride x 2 (dIx, dIy), zero border; NULL when not requested */
} OrcKltLevel;
/* levels: room for 8; returns the top level index (<= max_level). */
int orc_klt_build_levels(const uint8_t* img, int w, int h, int win, int max_level, int with_deriv, OrcKltLevel* levels);
void orc_klt_free_levels(OrcKltLevel* levels, int top);
void orc_pyrdown_std(const uint8_t* src, int w, int h, uint8_t* dst); /* cv::pyrDown -> ((w+1)/2,(h+1)/2) */
/* KLT::calc2D with flags = 0 (klt.cpp:491-628): prev_pts/next_pts n x 2 floats */
void orc_klt_calc2d(const uint8_t* img_ref, const uint8_t* img_cur, int w, int h, const float* prev_pts, int n,
                    int win, int max_level, int max_count, double epsilon, float min_eig, int illumination_adapt,
                    float* next_pts, uint8_t* status, float* err);
int orc_lm_anchors(const OrcLine* lines, int n_lines, int step, float* kps, int32_t* line_kp_num, int cap);
/* ref_to_cur: n_ref entries (-1 = unmatched).  Optional per-anchor outputs (cap_kp entries). */
int orc_line_matching(const uint8_t* img_ref, const uint8_t* img_cur, int w, int h, const OrcLine* lines_ref, int n_ref,
                      const OrcLine* lines_cur, int n_cur, const OrcLineMatchParam* P, int illumination_adapt,
                      int topological_filter, int32_t* ref_to_cur, float* kps_ref, float* kps_cur, uint8_t* status,
                      float* err, int32_t* kp2line_cur, int cap_kp, int* n_kp);

int64_t orc_linefront_sequence_mt(const uint8_t* frames, int n_frames, int w, int h, const OrcEDLineParam* p,
                                  int smoothed, int n_threads);

/* ---- vanishing-point stage after the path (orc_vp.c) ---- */
typedef struct { int32_t r[31]; int f, b; } OrcGRand; /* glibc rand() state, TYPE_3 */
void orc_grand_seed(OrcGRand* g, unsigned seed);      /* = srand(seed) */
int32_t orc_grand_next(OrcGRand* g);                  /* = rand()      */
#define ORC_VP_MAX_DRAWS 1000000
#define ORC_VP_FLAG_LX_OOB 1 /* the reference would have read lx[] out of range on this frame */
int orc_vp_hypothesis_count(void); /* "it" of getVPHypVia2Lines (105) */
/* vanishing_point_detection::run_vanishing_point_detection(img, lines, all_lines, vps, local_vp_ids)
 * after init(f, cx, cy, .).  seed: what time(NULL) returned; frame_count: calls made before this one.
 * math_mode 0: libm, 1: the shared deterministic functions.  vps: 9 doubles (3 unit vectors);
 * vp_idx: n_all labels 0..2, 3 = none.  Optional: grid (90 x 360 smoothed cells), best_idx
 * (hypothesis index), pairs (2 per outer iteration), flags, scores (the sum of every
 * hypothesis, 360 per outer iteration).  Returns 0, -1 (fewer than 2 lines), -2. */
int orc_vp_detect(const OrcLine* lines, int n_lines, const OrcLine* all_lines, int n_all, float f, float cx,
                  float cy, unsigned seed, int frame_count, int math_mode, double* vps, int32_t* vp_idx,
                  double* grid, int32_t* best_idx, int32_t* pairs, int32_t* flags, double* scores);
int64_t orc_vp_sequence(const OrcLine* lines, const int32_t* counts, int n_frames, int cap, float f, float cx, float cy,
                        const uint32_t* seeds, int frame_count0, int math_mode, double* vps, int32_t* vp_idx);
/* line_feature_tracker_node.cpp:64-153: cloud = 3n point floats then 7 channels of n floats (see orc_vp.c) */
void orc_line_cloud(const OrcLine* lines, const int32_t* ids, int n, const double* line_vps, int n_vps, float fx,
                    float fy, float cx, float cy, int num_of_cam, int cam, float* cloud);
double orc_atan2_cr(double y, double x);
double orc_atan_cr(double t);
double orc_acos_cr(double x);
double orc_sin_cr(double a);
void orc_sincos_cr(double a, double* s, double* c);

/* Whole front end on a frame sequence (for the CPU baseline timing).  Returns
 * total keylines over the sequence (each frame is matched k=1 against the
 * previous one; results are discarded, this entry point exists for timing). */
int64_t orc_frontend_sequence(const uint8_t* frames, int n_frames, int w, int h,
                              int num_octaves, int max_lines);

/* Same, frames split over n_threads host threads in contiguous chunks with a
 * one-frame halo (orc_threads.c). */
int64_t orc_frontend_sequence_mt(const uint8_t* frames, int n_frames, int w, int h,
                                 int num_octaves, int max_lines, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
