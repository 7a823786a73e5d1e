// vpl_common.cuh -- shared device/host declarations of libvplines_b200 (sm_100a).
//
// Data layout in HBM (one "slot" = one batch of B frames, per octave o):
//   img      B x H  x W      u8      raw frames (octave 0 only)
//   pyr[o]   B x Ho x Wo     u8      Gaussian pyramid (octave 0 = blur5(img) if blur_first)
//   grad[o]  B x Ho x Wo     short2  Sobel (dx,dy) of pyr[o]            (LBD)
//   scl[o]   B x Hs x Ws     u8      LSD working image: blur7 + 0.8 resize
//   ang[o]   B x Hs x Ws     f32     level-line angle in degrees ([0, 360]), -1024 = NOTDEF (read-only: NFA scans,
//                                    the speculative engine)
//   pix[o]   B x Hs x Ws     16 B    engine record {angle bits | USED bit31, cosf, sinf, packed gradient differences}:
//                                    ONE 128-bit gather per neighbour and one level of dependent loads per growth
//                                    step (measured: splitting it into 4-byte planes costs more sectors and more
//                                    load levels than it saves bytes)
//   ord[o]   B x Hs x Ws     i32     pseudo-ordered seed list (defined pixels only)
//   reg[o]   B x Hs x Ws     u32     region list of the frame's engine warp: packed pixel coordinates (x | y << 16);
//                                    the ordering kernel keeps its u16 bins there before the engine runs
//   (cos, sin) of a pixel's angle come from a 511 x 511 table indexed by the two differences (one per context,
//   2 MB, L2-resident), looked up once per pixel by ll_angle (no FP64 trigonometry per pixel); the region list
//   carries coordinates only (round 1: 16 bytes per entry)
//   cand[o]  B x cap         rects   post-refine rectangles in seed order
//   keylines B x cap x 68 B, desc B x cap x 32 B, matches B x cap x k
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vpl_capi.h"

namespace vpl {

constexpr int kMaxOctaves = 4;
constexpr int kBins = 1024;
constexpr float kNotDefDeg = -1024.0f;

constexpr uint32_t kUsedBit = 0x80000000u;   // ang plane: sign bit of a defined angle = used
// Per-pixel record of the default region engine.  ang: float bits of the level-line angle in degrees ([0,360]);
// NOTDEF is -1024.0f (bit31 set, so "not a candidate"); bit31 set on a defined pixel = USED.
// dabc = (DA + 255) | (BC + 255) << 16, the pixel's gradient differences (gx^2+gy^2 = modgrad^2 * 4).
struct __align__(16) Pix {
  uint32_t ang;
  float cs, sn;
  uint32_t dabc;
};
constexpr uint32_t kTagFree = 0xFFFFFFFFu;   // speculative engine's owner tags: a defined pixel no region owns
constexpr int kLutN = 511;                   // (cos, sin) table: [DA + 255][BC + 255], DA = d - a, BC = b - c of the 2x2 block

// gx^2 + gy^2 from the packed gradient differences (modgrad = sqrt(q / 4))
__device__ __forceinline__ int dabc_q(uint32_t w) {
  const int DA = (int)(w & 0xffffu) - 255, BC = (int)(w >> 16) - 255;
  const int gx = DA + BC, gy = DA - BC;
  return gx * gx + gy * gy;
}
// (cosf, sinf) of the level-line angle from the packed gradient differences
__device__ __forceinline__ float2 dabc_cssn(const float2* __restrict__ lut, uint32_t w) {
  return __ldg(lut + (w & 0xffffu) * kLutN + (w >> 16));
}

// A parked transaction of the region engine (finished, waiting for the commit pointer): where its pixel
// lists lie in the lane's ring, what it depends on, what to do at commit.
constexpr int kEngQ = 16;     // parked transactions per lane (power of two)
constexpr int kEngDeps = 4;   // recorded dependencies per transaction
struct __align__(16) EngDesc {
  int pos, start, ext, flags;
  int dep[kEngDeps];
  int nd, pad0, pad1, pad2;
};

// A rectangle candidate produced by the region engine (cv lsd.cpp `struct rect`).
struct __align__(16) RectCand {
  double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p;
  double nfa;      // filled by the NFA kernel
  int accepted;    // filled by the NFA kernel
  int pad;
};

// Per-(frame, octave) LSD geometry, constant over a batch.
struct OctaveGeom {
  int w, h;          // pyramid image size
  int ws, hs;        // 0.8-scaled size
  double log_nt;     // 5*(log10 ws + log10 hs)/2 + log10(11)
  int min_reg_size;  // int(-log_nt / log10(p))
};

struct LsdConst {
  double prec;  // pi * 22.5 / 180
  double p;     // 22.5 / 180
  double rho;   // 2 / sin(prec)
};

#define VPL_DEG2RAD (3.1415926535897932384626433832795 / 180.0)
#define VPL_PI 3.1415926535897932384626433832795
#define VPL_3_2_PI 4.71238898038468985769
#define VPL_2PI 6.28318530717958647692

__host__ __device__ inline int refl101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * n - 2 - i;
  }
  return i;
}

// cv::fastAtan2 scalar path: float32, explicit round-to-nearest mul/add so that
// nothing is contracted into an FMA (must equal oracle/orc_prims.c orc_fast_atan2).
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale;
  const float p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale;
  const float p7 = -0.04432655554792128f * scale;
  const float eps = (float)2.2204460492503131e-16;
  float ax = fabsf(x), ay = fabsf(y);
  // both branches of the scalar code are min / (max + eps) followed by the same polynomial: one copy, no divergence
  const bool steep = !(ax >= ay);
  const float c = __fdiv_rn(steep ? ax : ay, __fadd_rn(steep ? ay : ax, eps));
  const float c2 = __fmul_rn(c, c);
  float a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  if (steep) a = __fsub_rn(90.f, a);
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// Kernel launchers (host side, defined in the .cu files).
void launch_blur5_sobel(const uint8_t* img, uint8_t* pyr, short2* grad, int w, int h, int batch,
                        int do_blur, cudaStream_t st);
void launch_pyrdown(const uint8_t* src, uint8_t* dst, int w, int h, int batch, cudaStream_t st);
void launch_sobel(const uint8_t* src, short2* grad, int w, int h, int batch, cudaStream_t st);
void launch_scale08(const uint8_t* src, uint8_t* dst, int w, int h, int ws, int hs, int batch,
                    cudaStream_t st);
void launch_ll_angle(const uint8_t* scl, float* ang, Pix* pix, const float2* lut, unsigned int* maxq,
                     int ws, int hs, int batch, double rho, cudaStream_t st);
// (cosf, sinf) of the level-line angle for every pair of gradient differences: kLutN x kLutN float2
void launch_cssn_lut(float2* lut, cudaStream_t st);
// scratch: >= 2*ws*hs bytes per frame, frames `scratch_stride` bytes apart (the region
// scratch `reg` is free until the engine runs and is used for this)
void launch_order(const uint8_t* scl, const unsigned int* maxq, int* ord, int* n_ord, void* scratch,
                  size_t scratch_stride, int ws, int hs, int batch, double rho, cudaStream_t st);
struct EngineOct {
  Pix* pix;          // B x hs x ws engine records
  uint32_t* tag;     // speculative engine only: kSpecMaxBatch x hs x ws owner tags (0 = undefined, 0xFFFFFFFF = free)
  uint32_t* arena;   // speculative engine only: kSpecMaxBatch x 2 x hs x ws list entries
  EngDesc* desc;     // speculative engine only: kSpecMaxBatch x 32 x kEngQ
  RectCand* rects;   // speculative engine only: kSpecMaxBatch x 32 x kEngQ
  const float* ang;  // B x hs x ws level-line angle in degrees
  const int* ord;    // B x hs x ws
  const int* n_ord;  // B
  uint32_t* reg;     // B x hs x ws region list of the default engine
  RectCand* cand;    // B x cand_cap
  int* n_cand;       // B
  int ws, hs;
  double log_nt;
  int min_reg_size;
};
struct EngineArgs {
  EngineOct oct[kMaxOctaves];
  int num_octaves;
  int cand_cap;
  int batch;
  LsdConst lc;
  int* overflow;  // set to 1 if a frame produced more than cand_cap candidates
  const double* lgam;  // lgam[m] = log_gamma((double)m) for the NFA kernel
  int lgam_n;
  int ring_cap;        // list entries per lane of the speculative engine (0 = 2*ws*hs/32; tests force small values)
  const float2* lut;   // (cosf, sinf) table, kLutN x kLutN
};
// Two region engines with identical results (both bit-equal to the sequential CPU algorithm):
//   lsd_engine.cu (default): one warp per frame keeps the sequential seed order and spends its lanes on the
//     order-free work inside a step;
//   lsd_engine_spec.cu (opt-in, batches of at most kSpecMaxBatch frames): one warp per frame, 32 speculative seeds
//     in flight, committed in seed order.
void launch_region_engine(const EngineArgs& a, cudaStream_t st);
void launch_region_engine_spec(const EngineArgs& a, cudaStream_t st);
constexpr int kSpecMaxBatch = 256;
void launch_rect_nfa(const EngineArgs& a, cudaStream_t st);
void nfa_check_stats(unsigned long long out[4]);  // counters of the -DVPL_NFA_CHECK build, zeros otherwise
struct PackArgs {
  const RectCand* cand[kMaxOctaves];
  const int* n_cand[kMaxOctaves];
  int w[kMaxOctaves], h[kMaxOctaves];
  int num_octaves;
  int scale;
  int cand_cap;
};
void launch_pack_keylines(const PackArgs& a, VplKeyLine* kl, int* counts, int* overflow, int cap, int batch,
                          cudaStream_t st);
void launch_pack_segments(const RectCand* cand, const int* n_cand, int cand_cap, VplSegment* out, int* count,
                          int cap, int batch, cudaStream_t st);
struct LbdArgs {
  const short2* grad[kMaxOctaves];
  int w[kMaxOctaves], h[kMaxOctaves];
  int num_octaves;
};
void launch_lbd(const LbdArgs& a, const VplKeyLine* kl, const int* counts, int cap, uint8_t* desc, float* fdesc,
                int batch, cudaStream_t st);
void launch_hamming_knn(const uint8_t* q, const int* nq, int cap_q, const uint8_t* t, const int* nt, int cap_t,
                        int n_pairs, int k, VplDMatch* out, cudaStream_t st);
void lbd_init_tables();
void launch_remap(const uint8_t* src, uint8_t* dst, const float* mapx, const float* mapy, const uint16_t* wtab,
                  int w, int h, int batch, cudaStream_t st);
// lut: batch x tiles x tiles x 256 bytes of scratch
void launch_clahe(const uint8_t* src, uint8_t* dst, uint8_t* lut, int w, int h, double clip, int tiles, int batch,
                  cudaStream_t st);
// ---- EDLines (edlines.cu) ---------------------------------------------------------------------
// Per-frame geometry of one EDLines batch (all frames of a batch share it).
struct EdGeom {
  int w, h;
  int scan;            // scanIntervals_
  int nW, nH;          // scan-grid size: x = 1 + ix*scan < w-1, y = 1 + iy*scan < h-1
  int bm_words;        // 32-bit words of the column-major anchor bitmap
  int cap_px;          // w*h/5   (edgePixelArraySize, edline_detector.cpp:92)
  int cap_edges;       // cap_px/20 (maxNumOfEdge, :93)
  int part_cap;        // per-walk scratch entries
  int nslots;          // line slots per frame: 2*cap_px / minLineLen + 1
  int min_len;
  double logNT;        // 2*(log10 w + log10 h), :1186
};
// Device buffers of one slot (frame f's part of each array starts at f * the per-frame size
// that EdGeom gives for the batch's image size).
struct EdBuffers {
  uint16_t* gmap;      // B x h x w: bits 0-8 gradient/4, bit 13 edge mark, bit 15 horizontal
  unsigned* bitmap;    // B x bm_words
  int* n_anchor;       // B
  uint32_t* first;     // B x part_cap   x | y << 16 of the current chain's first part
  uint32_t* second;    // B x part_cap
  uint32_t* xy;        // B x 2*cap_px   edge chains, re-packed
  uint32_t* sid;       // B x (cap_edges + 2) chain starts
  int* n_chain;        // B
  int* n_px;           // B
  int* status;         // B: 1, or -1 where EdgeDrawing returns -1
  VplLine* slots;      // B x nslots
  uint8_t* slot_valid; // B x nslots
};
// img: the smoothed frames (B x h x w); writes grad (Sobel pair), B.gmap, B.bitmap, B.n_anchor
void launch_ed_grad_anchor(const uint8_t* img, short2* grad, const EdBuffers& B, const EdGeom& G, int grad_thresh,
                           int anchor_thresh, int batch, cudaStream_t st);
void launch_ed_walk(const EdBuffers& B, const EdGeom& G, int batch, cudaStream_t st);
void launch_ed_fit(const EdBuffers& B, const EdGeom& G, double fit_thr, const short2* grad, const double* lgam,
                   int batch, cudaStream_t st);
void launch_ed_compact(const EdBuffers& B, const EdGeom& G, VplLine* out, int* counts, int cap, int* overflow,
                       int batch, cudaStream_t st);
// ---- line matching (linematch.cu) ------------------------------------------------------------
constexpr int kKltMaxLevels = 4;  // maxLevel 3 (line_matching.cpp:14)
// Pyramid of one frame: every level stored with a `pad`-pixel border (REFLECT_101 for the image,
// zero for the Scharr pair), levels back to back.
struct KltGeom {
  int pad;                         // border rows above/below and valid border columns: the window size, 13
  int padx;                        // columns in front of pixel 0 (16: rows and pixel 0 stay 16-byte aligned)
  int top;                         // highest level built (cv::buildOpticalFlowPyramid's return)
  int w[kKltMaxLevels], h[kKltMaxLevels], stride[kKltMaxLevels];
  size_t img_off[kKltMaxLevels];   // bytes from the frame's pyramid base
  size_t deriv_off[kKltMaxLevels]; // short2 elements from the frame's derivative base
  size_t img_frame, deriv_frame;   // per-frame sizes (bytes / short2 elements)
};
struct LmParams {
  int step;
  float closest, ratio, dist_ratio, klt_err;
  int max_count;
  double eps2;
  float min_eig;
  float topo_dist, topo_len, topo_viol;
  int illum, topo;
};
struct LmBuffers {   // per pair p (cap_kp anchors, cap lines)
  int cap_kp;
  float2* kps;       // anchors on the reference lines
  float2* nxt;       // tracked positions
  uint8_t* status;
  float* err;
  int* kp2line;      // closest current line per anchor
  int* kp_start;     // (cap + 1) per pair: first anchor of each reference line
  int* n_kp;
  int* r2c;          // cap per pair: reference line -> current line
  int* matched;      // number of matches per pair
  int* pair_off;     // n_pairs + 1: exclusive prefix of n_kp (the tracker's work queue runs over all pairs)
  int* queue;        // kKltMaxLevels counters: next work item per pyramid level
  int* overflow;
};
void launch_klt_pyramid(const uint8_t* img, uint8_t* pyr, short2* deriv, const KltGeom& G, int w, int h, int batch,
                        cudaStream_t st);
void launch_lm_anchors(const VplLine* lines, const int* counts, int cap, const LmBuffers& B, const LmParams& P,
                       int pstride, int n_pairs, cudaStream_t st);
void launch_klt_track(const uint8_t* pyr, const short2* deriv, const KltGeom& G, const LmBuffers& B, const LmParams& P,
                      int pstride, int n_pairs, cudaStream_t st);
void launch_lm_vote(const VplLine* lines, const int* counts, int cap, const LmBuffers& B, const LmParams& P, int pstride,
                    int n_pairs, cudaStream_t st);
void launch_popc_peak(unsigned* out, int blocks, int iters, cudaStream_t st);
// ---- vanishing points (vp.cu) ---------------------------------------------------------------
struct VpParams {
  double f, ppx, ppy;   // vanishing_point_detection::init: floats stored in doubles
  int it;               // outer iterations of getVPHypVia2Lines (105)
  long long max_draws;  // give up (status -2) where the reference would redraw for ever
};
struct VpBuffers {      // per frame (cap lines)
  double* para;         // cap x 3: p1 x p2 per line; reused for the three angles per line of all_lines
  double* length;       // cap
  double* orient;       // cap; reused for segAngle of all_lines
  double* vp1;          // it x 3
  int* pairs;           // it x 2 (the line pair of every outer iteration)
  int* rng;             // 33: rand() state after the pair draws
  int* status;          // 0, -1 (fewer than 2 lines), -2
  int* flags;           // bit 0: the reference would have read lx[] out of range
  double* grid;         // 90 x 360 votes
  double* grid_new;     // 90 x 360 after the 3x3 pass
  double* lambda_sc;    // 360 x (sin, cos) of j * 2 pi / 360 (one table per context)
  double* part_best;    // kVpMaxSplits: best sum per CTA of the scoring kernel
  int* part_idx;
  int* best_idx;
  int* lx;              // cap: indices of the lines put into lx by lines2Vps
};
constexpr int kVpMaxSplits = 15;
constexpr int kVpCells = 90 * 360;
void launch_vp_lambda(double* lambda_sc, cudaStream_t st);
void launch_vp_cloud(const VplLine* all_lines, const int* n_all, int cap, const int* ids, const double* line_vps, float fx,
                     float fy, float cx, float cy, int num_of_cam, int cam, float* cloud, int n_frames, cudaStream_t st);
int vp_score_splits(int n_frames);
// the stage = these four, in this order (each: the launches it makes)
void launch_vp_prepare(const VplLine* lines, const int* n_lines, int cap, const unsigned* seeds, const VpBuffers& B,
                       const VpParams& P, int n_frames, cudaStream_t st);                       // memset + 1
void launch_vp_vote(const int* n_lines, int cap, const VpBuffers& B, const VpParams& P, int n_frames,
                    cudaStream_t st);                                                           // 2
void launch_vp_score(const VpBuffers& B, const VpParams& P, int n_frames, cudaStream_t st);      // 1
void launch_vp_scores_debug(const VpBuffers& B, const VpParams& P, int n_frames, int frame, double* scores, cudaStream_t st);
void launch_vp_classify(const VplLine* all_lines, const int* n_all, int cap, int frame_count0, const VpBuffers& B,
                        const VpParams& P, int n_frames, double* vps, int* vp_idx, double* line_vps,
                        cudaStream_t st);                                                       // 1
#ifdef VPL_DEBUG_NFA
void debug_set_cand(int c);
#endif

}  // namespace vpl
