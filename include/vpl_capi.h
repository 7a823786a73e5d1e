/*
 * vpl_capi.h -- C ABI of libvplines_b200.so: the B200-native line front end
 * (LSD line segments -> LBD binary descriptors -> brute-force Hamming kNN).
 *
 * This is the drop-in boundary of SURVEY.md section 8(b).  The reference links
 * its line primitives as a C++ static library (libline_matching.a,
 * /root/reference/feature_tracker/CMakeLists.txt:51-57) and reaches them through
 * two private seams of LineFeatureTracker::readImage:
 *     void edline_detect(cv::Mat&, std::vector<Line>&, const bool&)
 *     void match_line_match(const cv::Mat&, const cv::Mat&, std::vector<Line>&,
 *                           std::vector<Line>&, std::vector<int>&)
 *   (/root/reference/feature_tracker/include/linefeature_tracker.h:74-79, called
 *   at feature_tracker/src/line_feature_tracker.cpp:87 and :115).
 * The north star puts the OpenCV-3.4 line_descriptor surface at that position:
 *     LSDDetector::detect(const Mat&, vector<KeyLine>&, int scale, int numOctaves, const Mat& mask)
 *     BinaryDescriptor::compute(const Mat&, vector<KeyLine>&, Mat& descriptors, bool returnFloatDescr)
 *     BinaryDescriptorMatcher::match / knnMatch(const Mat& q, const Mat& t, ...)
 *   (opencv_contrib 3.4 modules/line_descriptor/include/opencv2/line_descriptor/
 *   descriptor.hpp; only the dead includes at feature_tracker/include/
 *   vanishing_point_detection.h:17-18 remain of it in the reference tree).
 * Each entry point below names the interface it replaces.  The header-only C++
 * facade vplines-slam_b200/compat/line_descriptor.hpp maps that surface (and the
 * Line / vector<int> seam, compat/vplines_seam.hpp) onto these calls.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function
 * returns 0 on success or a negative VPL_E_* code and never throws;
 * vpl_last_error() gives the message.  All device memory, pinned staging and
 * streams belong to the context; nothing is allocated on the per-batch path.
 * There is NO CPU fallback: without a CUDA device vpl_create fails.
 * A context is not re-entrant (same as the reference's detector/matcher objects,
 * SURVEY.md 8b "Threading"); use one context per host thread per GPU.
 */
#ifndef VPL_CAPI_H
#define VPL_CAPI_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPL_OK 0
#define VPL_E_INVALID (-1)  /* bad argument                                   */
#define VPL_E_CUDA (-2)     /* CUDA runtime error (message has the detail)     */
#define VPL_E_CAPACITY (-3) /* batch/line/image larger than the context allows */
#define VPL_E_NODEVICE (-4) /* no usable CUDA device: there is no CPU path     */

/* cv::line_descriptor::KeyLine, field for field (17 x 4 bytes, no padding). */
typedef struct VplKeyLine {
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
} VplKeyLine;

/* cv::DMatch */
typedef struct VplDMatch {
  int32_t queryIdx, trainIdx, imgIdx;
  float distance;
} VplDMatch;

/* One raw LSD segment (cv::LineSegmentDetector::detect outputs _lines, width,
 * prec, nfa). */
typedef struct VplSegment {
  float x1, y1, x2, y2;
  double width, prec, nfa;
} VplSegment;

/* The numeric fields of the reference's struct Line (line_matching/src/line.h:8-12): endpoints
 * in pixels, unit-normal line equation w1 x + w2 y + w3 = 0, centre, length.  56 bytes. */
typedef struct VplLine {
  float endpoint[4];
  double equation[3];
  float center[2];
  float length;
  int32_t reserved; /* 0 */
} VplLine;

/* EDLineParam, field for field (line_matching/src/edline_detector.h:32-40). */
typedef struct VplEDLineParam {
  int32_t ksize;            /* Gaussian kernel size, only used when smoothed == 0 (5 supported) */
  float sigma;              /* ... and its sigma (1.0 supported)                                */
  float gradientThreshold;  /* stored as short by the detector (edline_detector.h:118)          */
  float anchorThreshold;    /* stored as unsigned char (:122)                                   */
  int32_t scanIntervals;
  int32_t minLineLen;
  double lineFitErrThreshold;
} VplEDLineParam;

typedef struct VplContext VplContext;

typedef struct VplConfig {
  int32_t device;       /* CUDA device ordinal                                   */
  int32_t max_width;    /* largest image the context accepts                     */
  int32_t max_height;
  int32_t max_octaves;  /* >= numOctaves of any call (1..4)                      */
  int32_t max_lines;    /* KeyLine capacity per frame (all octaves together)     */
  int32_t max_batch;    /* frames per batch                                      */
  int32_t num_slots;    /* batches in flight for submit/collect (1..4)           */
  int32_t blur_first;   /* 1: LSDDetector blurs octave 0 with GaussianBlur 5x5   */
  int32_t profile;      /* 1: record per-stage CUDA-event timings                */
  int32_t lsd_path;     /* 1 (default): allocate the LSD/LBD/Hamming buffers; 0: a context for the
                           EDLines + KLT line matching front end only (saves ~13 MB per frame)    */
} VplConfig;

/* ---- lifetime -------------------------------------------------------------- */
void vpl_default_config(VplConfig* cfg);
int vpl_create(const VplConfig* cfg, VplContext** out);
void vpl_destroy(VplContext* ctx);
const char* vpl_last_error(const VplContext* ctx); /* ctx may be NULL: last create error */
const char* vpl_version(void);
int vpl_device_count(void);

/* ---- pre-processing in front of the path ----------------------- */
/* Replaces cv::remap(_img, img, undist_map1_, undist_map2_, CV_INTER_LINEAR) and
 * cv::createCLAHE(3.0, Size(8,8))->apply (LineFeatureTracker::readImage,
 * feature_tracker/src/line_feature_tracker.cpp:62, :64-68).  mapx/mapy: the CV_32FC1 maps of
 * initUndistortRectifyMap (w x h floats each; both NULL = no remap).  clahe_clip <= 0 = no CLAHE
 * (the reference uses 3.0 with an 8x8 grid when EQUALIZE is set).  Once set, every call that takes
 * frames (detect / compute / frontend) applies it to the uploaded frames on the device first. */
int vpl_set_preprocess(VplContext* ctx, const float* mapx, const float* mapy, int w, int h,
                       double clahe_clip, int clahe_tiles);
/* The pre-processed frames themselves (n*w*h bytes), for the parity tests. */
int vpl_preprocess_batch(VplContext* ctx, const uint8_t* const* imgs, int n, int w, int h,
                         size_t stride, uint8_t* out);

/* ---- EDLineDetector::EDline: the detector the reference really runs ------ */
/* vpl_edlines_configure = EDLineDetector(EDLineParam) (line_matching/src/edline_detector.cpp:30-40;
 * the tracker node builds it with {5, 1.0, 30, 5, 2, min_line_length, line_fit_err},
 * feature_tracker/src/line_feature_tracker_node.cpp:203).  Allocates the detector's device
 * buffers for the context's max image size and batch; call it before the first detect and not
 * while batches are in flight. */
void vpl_edlines_default_param(VplEDLineParam* p);
int vpl_edlines_configure(VplContext* ctx, const VplEDLineParam* p);
/* = int EDLineDetector::EDline(cv::Mat& image, std::vector<Line>& lines, bool smoothed)
 * (edline_detector.cpp:1176-1198), i.e. the tracker's edline_detect seam
 * (feature_tracker/src/line_feature_tracker.cpp:315-321), on n frames.  lines: n * cap entries,
 * frame f at lines + f*cap, counts[f] valid, in (edge chain, position) order = the reference's
 * single-thread order (its multi-thread order is unspecified).  status (may be NULL): per frame 1,
 * or -1 where EdgeDrawing returns -1 (no anchors / capacity limits W*H/5, W*H/100,
 * edline_detector.cpp:166-169, :655-670) -- such a frame yields 0 lines.  More than cap lines in a
 * frame -> VPL_E_CAPACITY.  smoothed == 0 blurs with GaussianBlur(ksize 5, sigma 1) first. */
int vpl_edlines_detect_batch(VplContext* ctx, const uint8_t* const* imgs, int n, int w, int h,
                             size_t stride, int smoothed, VplLine* lines, int32_t* counts, int cap,
                             int32_t* status);
/* Pipelined form (same slot rules as vpl_frontend_submit/collect). */
int vpl_edlines_submit(VplContext* ctx, int slot, const uint8_t* const* imgs, int n, int w, int h,
                       size_t stride, int smoothed);
int vpl_edlines_collect(VplContext* ctx, int slot, VplLine* lines, int32_t* counts, int cap,
                        int32_t* status);
/* Re-runs the detector on the frames already resident in slot's device input buffer (kernel
 * timing); results stay in HBM; does not synchronise. */
int vpl_edlines_run_resident(VplContext* ctx, int slot);
/* Edge chains of frame `frame` of the last batch on slot 0 (EdgeChains, edline_detector.h:17-22):
 * xy = x | y << 16 per edge pixel (cap_px entries), sid = chain starts (cap_chains + 1 entries). */
int vpl_debug_edge_chains(VplContext* ctx, int frame, uint32_t* xy, int cap_px, uint32_t* sid,
                          int cap_chains, int32_t* n_px, int32_t* n_chains);

/* ---- LineMatching::Matching: the matcher the reference really runs ------- */
/* LineMatching's constructor defaults (line_matching/src/line_matching.h:14-18), the KLT it builds
 * and reconfigures in Matching() (line_matching.cpp:14, :630-631: window 13x13 -- fixed here --,
 * maxLevel 3, 30 iterations, eps 0.001, minEig 1e-4, flags 0), TopologicalFilter's defaults
 * (line_matching.h:45-47) and the two switches the tracker passes as true
 * (feature_tracker/src/line_feature_tracker.cpp:307-308). */
typedef struct VplLineMatchParam {
  int32_t step;
  float closest_line_threshold, line_matching_ratio, line_distance_error_ratio, klt_error_threshold;
  int32_t max_level, max_count;
  double epsilon;
  float min_eig;
  float topo_distance_threshold, topo_length_ratio, topo_violation_ratio;
  int32_t illumination_adapt, topological_filter;
  int32_t max_anchors; /* capacity: anchor points per frame pair (sum over lines of len/step + 2) */
} VplLineMatchParam;
void vpl_linematch_default_param(VplLineMatchParam* p);
/* = LineMatching(...) + its KLT: allocates the pyramids and anchor buffers for the context's max
 * image size and batch.  Call before the first match, not while batches are in flight. */
int vpl_linematch_configure(VplContext* ctx, const VplLineMatchParam* p);
/* = bool LineMatching::Matching(img_ref, img_cur, lines_ref, lines_cur, line_ref_to_line_cur, NULL,
 * NULL, NULL, illumination_adapt, topological_filter, 0) (line_matching.cpp:605-690), i.e. the
 * tracker's match_line_match seam (feature_tracker/src/line_feature_tracker.cpp:291-313), on n_pairs
 * independent pairs.  lines_ref: n_pairs * cap entries (pair p at + p*cap, n_ref[p] valid), lines_cur
 * likewise; ref_to_cur: n_pairs * cap, entry i of pair p = index of the current line matched to
 * reference line i, or -1 (all -1 when a side is empty, where Matching returns false).
 * n_pairs <= max_batch / 2. */
int vpl_linematch_batch(VplContext* ctx, const uint8_t* const* imgs_ref, const uint8_t* const* imgs_cur,
                        int n_pairs, int w, int h, size_t stride, const VplLine* lines_ref,
                        const int32_t* n_ref, const VplLine* lines_cur, const int32_t* n_cur, int cap,
                        int32_t* ref_to_cur);
/* Per-anchor results of pair `pair` of the last match on slot 0 (LineMatching::getPointMatchResult
 * + status_/errors_): cap entries each; *n = number of anchors. */
int vpl_debug_linematch_points(VplContext* ctx, int pair, float* kps_ref, float* kps_cur,
                               uint8_t* status, float* err, int32_t* kp2line_cur, int cap, int32_t* n);

/* ---- the reference's per-frame hot loop, fused: EDline on every frame + Matching(frame f-1,
 *      frame f) (LineFeatureTracker::readImage, line_feature_tracker.cpp:87 and :115) --------- */
/* n consecutive frames; lines/counts as vpl_edlines_detect_batch; prev_to_cur: n * cap, row f
 * (f >= 1) maps the lines of frame f-1 to the lines of frame f (-1 = unmatched; row 0 is all -1:
 * a caller that cuts a sequence into batches overlaps them by one frame).  Needs both
 * vpl_edlines_configure and vpl_linematch_configure. */
int vpl_linefront_batch(VplContext* ctx, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                        int smoothed, VplLine* lines, int32_t* counts, int cap, int32_t* prev_to_cur);
int vpl_linefront_submit(VplContext* ctx, int slot, const uint8_t* const* imgs, int n, int w, int h,
                         size_t stride, int smoothed);
int vpl_linefront_collect(VplContext* ctx, int slot, VplLine* lines, int32_t* counts, int cap,
                          int32_t* prev_to_cur);
int vpl_linefront_run_resident(VplContext* ctx, int slot);

/* ---- vanishing_point_detection::run_vanishing_point_detection: the stage the tracker runs on the
 *      matched lines of every frame (feature_tracker/src/vanishing_point_detection.cpp:37-65,
 *      called at feature_tracker/src/line_feature_tracker.cpp:241 / :243) ------------------------ */
/* = vanishing_point_detection::init(f, cx, cy, noiseRatio) (vanishing_point_detection.cpp:29-34; the
 * noise ratio argument is shadowed by a local 0.5 at :93 and never read).  Allocates the sphere grids
 * (2 x 259 KB per frame) and line scratch for max_batch frames of max_lines lines.  Call before the
 * first detection, not while batches are in flight. */
int vpl_vp_configure(VplContext* ctx, float f, float cx, float cy);
/* n_frames independent calls of run_vanishing_point_detection(img, lines, all_lines, vps, local_vp_ids)
 * (the image argument is only drawn on).  lines: n_frames * cap entries (frame i at + i*cap, n_lines[i]
 * valid) -- the set the hypotheses and the sphere vote use; all_lines / n_all: the set that is
 * classified (NULL = the same set: what readImage passes unless it found > 2 vertical lines).
 * seeds[i]: what time(NULL) returned for frame i -- the reference seeds rand() with it on every call
 * (:107).  frame_count0: calls the object made before frame 0 (frame i runs with frame_count0 + i; only
 * 0 / not 0 matters, :337-349).  Outputs: vps n_frames * 9 doubles (three unit vectors), vp_idx
 * n_frames * cap labels (0..2, 3 = none), optional line_vps n_frames * cap * 4 doubles (the Vector4d
 * readImage stores per line, line_feature_tracker.cpp:246-262), optional status per frame:
 * 0 ok, 1 ok but the reference itself would have read lx[] out of range on this frame (oracle/orc_vp.c),
 * -1 fewer than 2 lines (nothing labelled, vps zero: readImage's "no vp lines" branch), -2 no
 * non-degenerate line pair found. */
int vpl_vp_detect_batch(VplContext* ctx, const VplLine* lines, const int32_t* n_lines, const VplLine* all_lines,
                        const int32_t* n_all, int n_frames, int cap, const uint32_t* seeds, int frame_count0,
                        double* vps, int32_t* vp_idx, double* line_vps, int32_t* status);
/* The same, pipelined over the context's slots. */
int vpl_vp_submit(VplContext* ctx, int slot, const VplLine* lines, const int32_t* n_lines, const VplLine* all_lines,
                  const int32_t* n_all, int n_frames, int cap, const uint32_t* seeds, int frame_count0);
int vpl_vp_collect(VplContext* ctx, int slot, int cap, double* vps, int32_t* vp_idx, double* line_vps,
                   int32_t* status);
/* The sensor_msgs::PointCloud body img_callback publishes for every frame of the batch last collected from
 * `slot` (feature_tracker/src/line_feature_tracker_node.cpp:64-153), for camera index `cam` of num_of_cam
 * (NUM_OF_CAM; the loop variable i of :88): frame i's block starts at cloud + i*cap*10 and holds, for its
 * n = n_all[i] lines, 3n point floats -- ((x1 - cx)/fx, (y1 - cy)/fy, 1): the first endpoint through
 * LineFeatureTracker::undistortedLineEndPoints, line_feature_tracker.cpp:36-52 -- then the seven channels
 * id_of_line (= line_ids * num_of_cam + cam), u_of_endpoint, v_of_endpoint (second endpoint), vp_x, vp_y, vp_z,
 * vp_z_inv, n floats each.  As in the reference, every line carries vp[cam]: the Vector4d of line number `cam`
 * (:108-114 index the per-line list with the camera index).  line_ids: n_frames * cap (the tracker's lineID). */
int vpl_vp_pack_cloud(VplContext* ctx, int slot, const int32_t* line_ids, int cap, float fx, float fy, float cx, float cy,
                      int num_of_cam, int cam, float* cloud);
/* Re-runs the stage on the lines already resident on the slot (measurement). */
int vpl_vp_run_resident(VplContext* ctx, int slot);
/* Stage outputs of frame `frame` of the last batch on slot 0: the smoothed 90 x 360 grid, the index of
 * the winning hypothesis, the line pair of every outer iteration (2 x 105 ints).  Any may be NULL. */
int vpl_debug_vp(VplContext* ctx, int frame, double* grid, int32_t* best_idx, int32_t* pairs);
/* The score (sum of its three cells) of every hypothesis of that frame: 105 * 360 doubles, index = outer iteration *
 * 360 + angle index (lineLength[] of getBestVpsHyp, vanishing_point_detection.cpp:285-318). */
int vpl_debug_vp_scores(VplContext* ctx, int frame, double* scores);

/* ---- LineFeatureTracker::readImage's line pipeline, fused, for n consecutive frames in one pass over the device
 *      (feature_tracker/src/line_feature_tracker.cpp:56-288): remap + CLAHE (:62-68, when vpl_set_preprocess
 *      configured them) -> EDline (:87) -> Matching(frame f-1, frame f) (:115) -> run_vanishing_point_detection on
 *      each frame's own detected lines (lines == all_lines, what :243 passes on a frame whose lines are all new; the
 *      tracker's bookkeeping of tracked / new lines between :120 and :230 is host logic and stays with the caller).
 *      The lines never leave the device between the stages.  Needs vpl_edlines_configure, vpl_linematch_configure
 *      and vpl_vp_configure.  seeds / frame_count0 as vpl_vp_detect_batch; the other arguments and outputs as
 *      vpl_linefront_* and vpl_vp_collect (vp_idx: n * cap labels, one per detected line). -------------------- */
int vpl_readimage_submit(VplContext* ctx, int slot, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                         int smoothed, const uint32_t* seeds, int frame_count0);
int vpl_readimage_collect(VplContext* ctx, int slot, VplLine* lines, int32_t* counts, int cap, int32_t* prev_to_cur,
                          double* vps, int32_t* vp_idx, int32_t* vp_status);
/* Re-runs the whole pipeline on the frames resident on the slot (from the raw frames when pre-processing is on). */
int vpl_readimage_run_resident(VplContext* ctx, int slot);

/* ---- LSDDetector::detect (replaces edline_detect, linefeature_tracker.h:74) -- */
/* imgs: n host pointers to CV_8UC1 images of w x h with row pitch `stride` bytes.
 * keylines: n * cap entries, frame f at keylines + f*cap; counts[f] = number found
 * (<= cap; if a frame has more than cap lines the call returns VPL_E_CAPACITY). */
int vpl_lsd_detect_batch(VplContext* ctx, const uint8_t* const* imgs, int n, int w, int h,
                         size_t stride, int scale, int num_octaves, VplKeyLine* keylines,
                         int32_t* counts, int cap);

/* ---- BinaryDescriptor::compute --------------------------------------------- */
/* keylines/counts laid out as above (caller-supplied, any detector). desc: n*cap*32
 * bytes, row i of frame f at desc + (f*cap+i)*32 (CV_8UC1 rows of 32). */
int vpl_lbd_compute_batch(VplContext* ctx, const uint8_t* const* imgs, int n, int w, int h,
                          size_t stride, const VplKeyLine* keylines, const int32_t* counts,
                          int cap, uint8_t* desc);

/* Same with returnFloatDescr = true: fdesc holds n*cap*72 floats (CV_32FC1 rows of 72). */
int vpl_lbd_compute_float_batch(VplContext* ctx, const uint8_t* const* imgs, int n, int w, int h,
                                size_t stride, const VplKeyLine* keylines, const int32_t* counts,
                                int cap, float* fdesc);

/* ---- BinaryDescriptorMatcher::match / knnMatch (replaces match_line_match,
 *      linefeature_tracker.h:76-79) ------------------------------------------- */
/* n_pairs independent (query, train) problems.  q: n_pairs*cap_q*32 bytes, pair p at
 * q + p*cap_q*32 with nq[p] valid rows; t likewise.  Brute force, ascending
 * distance, lowest train index on ties.  out: n_pairs*cap_q*k DMatch; entries
 * beyond nt[p] candidates have trainIdx=-1. */
int vpl_match_batch(VplContext* ctx, const uint8_t* q, const int32_t* nq, int cap_q,
                    const uint8_t* t, const int32_t* nt, int cap_t, int n_pairs, int k,
                    VplDMatch* out);

/* Re-runs the matcher on the descriptor sets left resident by the last vpl_match_batch (measurement). */
int vpl_match_run_resident(VplContext* ctx, int k);
/* POPC throughput of the device in popc32 per second, measured with a micro-benchmark kernel (8 independent
 * popc chains per thread, 8 resident CTAs of 256 threads per SM): the denominator for the matcher's
 * integer-pipe utilisation. */
int vpl_debug_popc_peak(VplContext* ctx, double* popc32_per_s);

/* ---- the fused path: detect + compute + match(t, t-1), one batch ----------- */
/* Equivalent to the three calls above on n consecutive frames with everything
 * kept in HBM in between.  matches: n*cap*k entries; frame f (f>=1) is matched
 * (query) against frame f-1 (train); frame 0 is matched against the last frame of
 * the previous batch on this context if `chain` is non-zero, otherwise it gets
 * trainIdx=-1.  Any of keylines/desc/matches may be NULL to skip its download. */
int vpl_frontend_batch(VplContext* ctx, const uint8_t* const* imgs, int n, int w, int h,
                       size_t stride, int scale, int num_octaves, int k, int chain,
                       VplKeyLine* keylines, int32_t* counts, int cap, uint8_t* desc,
                       VplDMatch* matches);

/* Pipelined form: submit() uploads and enqueues batch work on slot s and returns
 * without waiting; collect() waits for slot s and downloads.  Up to num_slots
 * batches may be in flight. imgs must stay valid until submit returns (they are
 * staged into pinned memory inside submit). */
int vpl_frontend_submit(VplContext* ctx, int slot, const uint8_t* const* imgs, int n, int w, int h,
                        size_t stride, int scale, int num_octaves, int k, int chain);
int vpl_frontend_collect(VplContext* ctx, int slot, VplKeyLine* keylines, int32_t* counts, int cap,
                         uint8_t* desc, VplDMatch* matches);
/* Two batches per slot: a batch staged with vpl_frontend_upload may be submitted (imgs == NULL, or
 * vpl_frontend_submit_group) while the slot's previous front-end batch has not been collected yet.  Its kernels queue
 * behind the previous batch's on the slot's stream and its results go to the slot's second result generation, so
 * collecting the previous batch -- waiting for it, downloading its rows -- overlaps kernels.  Collects return the
 * batches of a slot oldest first.  At most two uncollected batches per slot; the next vpl_frontend_upload on the slot
 * needs the older one collected (it overwrites that batch's input buffer).  With per-stage events on
 * (VplConfig.profile) a submit waits for the slot's previous batch, which serialises the two.
 *
 * Upload ahead: copies the NEXT batch of slot s to the device on a copy stream of its own, into a
 * second input buffer of the slot, and returns without waiting -- allowed while the slot's current
 * batch is in flight, so the host-to-device copy of batch i+num_slots overlaps the kernels of batch
 * i instead of following its collect.  The next vpl_frontend_submit on that slot takes these frames
 * when it is called with imgs == NULL (n, w, h as uploaded).  The frames have to be contiguous
 * (stride == w, imgs[f] == imgs[0] + f*w*h), inside a vpl_host_register range, and left untouched
 * until that batch is collected.  (The reference has no counterpart: its readImage,
 * feature_tracker/src/line_feature_tracker.cpp:52, is handed one decoded frame at a time.) */
int vpl_frontend_upload(VplContext* ctx, int slot, const uint8_t* const* imgs, int n, int w, int h,
                        size_t stride);
/* Group form of vpl_frontend_submit for batches staged with vpl_frontend_upload: the batches of
 * slots[0..n_slots) (n[i] frames each, all w x h) are enqueued together, each on its slot's stream,
 * with a device-side barrier across the group in front of the region engine -- the engine launches
 * of the group then start together and fill the SMs' warp slots between them (the engine is bound
 * by the latency of one frame's sequential growth and wants 64 warps per SM; next to another
 * slot's streaming kernels both lose).  chain[i] as in vpl_frontend_submit; slot slots[i] is
 * matched against slots[i-1] of the group.  Each slot is collected on its own
 * (vpl_frontend_collect / _collect_dense).  Nothing is enqueued if any slot is not ready. */
int vpl_frontend_submit_group(VplContext* ctx, int n_slots, const int* slots, const int* n, int w,
                              int h, int scale, int num_octaves, int k, const int* chain);

/* Dense form of collect: the batch's KeyLines / descriptors / matches come back packed frame
 * after frame (frame f's rows start at sum(counts[0..f))), cap_total rows of capacity,
 * *total = rows written.  Output buffers registered with vpl_host_register are written by
 * the device-to-host copy itself (no staging copy). */
int vpl_frontend_collect_dense(VplContext* ctx, int slot, int32_t* counts, VplKeyLine* keylines,
                               uint8_t* desc, VplDMatch* matches, int64_t cap_total, int64_t* total);

/* Optional: pin a host buffer that holds frames (cudaHostRegister).  A submit whose frames
 * are contiguous (stride == w, imgs[f] == imgs[0] + f*w*h) and lie inside a registered range
 * is uploaded straight from it, asynchronously, without the staging copy -- the caller must
 * then leave those frames untouched until the slot is collected. */
int vpl_host_register(VplContext* ctx, const void* ptr, size_t bytes);
int vpl_host_unregister(VplContext* ctx, const void* ptr);
/* Bytes the last collect on `slot` copied device -> host. */
int64_t vpl_last_d2h_bytes(const VplContext* ctx, int slot);

/* Device-resident form used to time the kernels alone: runs the fused path on
 * n frames already in the context's device input buffer of slot s (filled by the
 * last submit on that slot), leaves results in HBM, does not synchronise. */
int vpl_frontend_run_resident(VplContext* ctx, int slot, int k);
/* The same for a group of slots, with the barrier of vpl_frontend_submit_group. */
int vpl_frontend_run_resident_group(VplContext* ctx, int n_slots, const int* slots, int k);
int vpl_sync(VplContext* ctx);

/* ---- raw stages, exported for the parity tests ------------------------------ */
/* cv::LineSegmentDetector(LSD_REFINE_ADV)::detect on one image (no pyramid blur). */
int vpl_lsd_raw(VplContext* ctx, const uint8_t* img, int w, int h, size_t stride, VplSegment* out,
                int32_t* count, int cap);
/* Image primitives on one image: which = 0 GaussianBlur5x5 s1 (u8), 1 pyrDown half
 * (u8, (w/2)x(h/2)), 2 Sobel dx,dy (int16 interleaved, 2*w*h), 3 LSD 0.8 scaling
 * (u8, out_w x out_h), 4 level-line angle in degrees (float, <0 undefined) of the
 * 0.8-scaled image, 5 pseudo-ordered pixel indices of the 0.8-scaled image (int32;
 * out_w = count). out must hold the result; out_w/out_h receive its size. */
int vpl_debug_stage(VplContext* ctx, int which, const uint8_t* img, int w, int h, size_t stride,
                    void* out, size_t out_bytes, int32_t* out_w, int32_t* out_h);

/* Candidate rectangles of the last vpl_lsd_raw call (post-refine rectangles in seed order;
 * after NFA validation the geometry is the improved one): 16 doubles each =
 * x1 y1 x2 y2 width x y theta dx dy prec p nfa accepted 0 0.  count receives the number. */
int vpl_debug_candidates(VplContext* ctx, double* out, int32_t* count, int cap);

/* ---- measurement ------------------------------------------------------------ */
#define VPL_STAGE_H2D 0
#define VPL_STAGE_PYRAMID 1   /* blur5+sobel, pyrDown+sobel          */
#define VPL_STAGE_SCALE 2     /* LSD 7x7 blur + 0.8 resize           */
#define VPL_STAGE_ANGLE 3     /* gradient / level-line angle         */
#define VPL_STAGE_ORDER 4     /* pseudo-ordering                     */
#define VPL_STAGE_REGION 5    /* region growing + rect + refine      */
#define VPL_STAGE_NFA 6       /* rectangle NFA validation            */
#define VPL_STAGE_PACK 7      /* compaction + KeyLine packing        */
#define VPL_STAGE_LBD 8
#define VPL_STAGE_MATCH 9
#define VPL_STAGE_D2H 10
#define VPL_STAGE_PREPROC 11  /* remap + CLAHE (optional)           */
#define VPL_STAGE_ED_GRAD 12    /* EDLines: Sobel pair, gradient/direction map, anchor bitmap */
#define VPL_STAGE_RESERVED13 13  /* reserved (keeps the stage numbering stable)                    */
#define VPL_STAGE_ED_WALK 14    /* smart routing (edge chains)     */
#define VPL_STAGE_ED_FIT 15     /* line fit + validation + compaction */
#define VPL_STAGE_LM_PYRAMID 16 /* line matching: KLT pyramids + Scharr */
#define VPL_STAGE_LM_TRACK 17   /* anchors + pyramidal LK               */
#define VPL_STAGE_LM_VOTE 18    /* closest line, vote, topological filter */
#define VPL_STAGE_VP_PREP 19     /* vanishing points: line parameters, rand() pairs (+ grid clear) */
#define VPL_STAGE_VP_VOTE 20     /* sphere-grid vote + 3x3 pass                                   */
#define VPL_STAGE_VP_SCORE 21    /* 105 x 360 hypotheses built and scored                         */
#define VPL_STAGE_VP_CLASSIFY 22 /* best hypothesis, line classification                          */
#define VPL_NUM_STAGES 23
/* Accumulated device milliseconds and launch counts per stage since the last
 * reset (cfg.profile must be 1).  ms/launches: arrays of VPL_NUM_STAGES. */
int vpl_get_stage_times(VplContext* ctx, double* ms, int64_t* launches);
int vpl_reset_stage_times(VplContext* ctx);
/* Where the stages of a slot's last batch lie in time (profile on): vpl_debug_mark records the origin on slot 0's
 * stream, vpl_debug_timeline waits for `slot` and returns start / end of each of its VPL_NUM_STAGES stages in ms after
 * the origin (-1 for a stage that did not run).  How the batches of two slots overlap on the device: bench.py --timeline. */
/* Counters of a library built with -DVPL_NFA_CHECK (zeros otherwise): early-exit decisions in the binomial tail of nfa(),
 * how many the float32 test answered, how many of those disagree with the double sequence (lsd_nfa.cu nfa_exit_fast). */
int vpl_debug_nfa_stats(VplContext* ctx, uint64_t* out4);
int vpl_debug_mark(VplContext* ctx);
int vpl_debug_timeline(VplContext* ctx, int slot, double* start_ms, double* end_ms);
/* Test hook: list entries per lane of the LSD region engine's rings (0 = default, 2*ws*hs/32 rounded down to a power
 * of two).  Small values force the engine's fallbacks (undo of parked regions, whole-arena mode); results must not change. */
int vpl_debug_set_engine_ring_cap(VplContext* ctx, int entries_per_lane);
/* Which LSD region engine runs.  0 = the default (one warp per frame in the sequential seed order), 1 = the
 * speculative engine (one warp per frame, 32 seeds in flight, committed in seed order; batches of at most 1024
 * frames, larger ones use the default).  Both produce the sequential algorithm's results bit for bit. */
int vpl_debug_set_engine(VplContext* ctx, int kind);
/* Switch the per-stage event timing on or off (VplConfig.profile) between batches; waits for the slots' streams.
 * With profiling on, a submit on a slot first waits for that slot's previous batch (its events are re-recorded). */
int vpl_set_profile(VplContext* ctx, int on);
/* Total kernels launched by this context since creation. */
int64_t vpl_kernel_launches(const VplContext* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VPL_CAPI_H */
