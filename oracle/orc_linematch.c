/*
 * orc_linematch.c -- CPU ORACLE (test infrastructure only, see vpl_oracle.h) for the reference's
 * real line matcher, SURVEY.md 8(f)-2: LineMatching::Matching
 *   /root/reference/line_matching/src/line_matching.cpp          (lm.cpp:line)
 *   /root/reference/line_matching/src/klt.cpp                    (klt.cpp:line)   KLT::calc2D
 *   /root/reference/line_matching/src/lk_tracker_invoker_2d.cpp  (lk2d.cpp:line)  per-point tracker
 * called by the tracker at feature_tracker/src/line_feature_tracker.cpp:115 -> :291-313 with
 * illumination_adapt = true, topological_filter = true, no K / T matrices (affines == nullptr).
 *
 * A restatement: anchors sampled on each reference line -> pyramidal Lucas-Kanade with
 * gain/bias adaptation -> closest current line per tracked anchor -> vote per reference line
 * -> topological (sidedness) filter.
 * PINNED: tests/test_oracle_linematch.py compares every output bit for bit with the reference's own
 * lk_tracker_invoker_2d.cpp and line_matching.cpp compiled against oracle/cvshim
 * (oracle/_ref/libref_linefront.so) and with the goldens that build produced
 * (tests/golden/ref_linematch.npz).  OpenCV library calls on the path are restated with the
 * semantics probed on cv2 4.13: buildOpticalFlowPyramid (pyrDown to (w+1)/2, REFLECT_101 padding of
 * winSize, early stop when a level would be <= winSize), the Scharr derivative of
 * KLT::calcSharrDeriv (= cv2.Scharr, REFLECT_101), meanStdDev on CV_16S (integer sum, double sum of
 * squares, mean = s * (1/N), sd = sqrt(max(q * (1/N) - mean^2, 0))), cvFloor / cvRound (SSE:
 * nearest-even, INT_MIN on NaN/overflow).
 *
 * All float arithmetic is sequential single precision exactly as the reference writes it (no FMA:
 * the oracle is built with -ffp-contract=off, the reference with plain -O3 on x86-64).
 */
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "vpl_oracle.h"

#define W_BITS 14
#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n)) /* CV_DESCALE, klt.h:39 */

static int cv_floor(float v) { /* cvFloor(float), SSE path: INT_MIN on NaN / out of range */
  if (!(v > -2147483648.0f && v < 2147483648.0f)) return INT_MIN;
  int i = (int)v;
  return i - (v < (float)i);
}
static int cv_round(float v) { /* cvRound(float): _mm_cvtss_si32, nearest even */
  if (!(v > -2147483648.0f && v < 2147483648.0f)) return INT_MIN;
  return (int)lrintf(v);
}
static int refl101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

/* ---- pyramid (cv::buildOpticalFlowPyramid, called at klt.cpp:590-594) ---------------------- */
void orc_pyrdown_std(const uint8_t* src, int w, int h, uint8_t* dst) { /* cv::pyrDown, default size */
  static const int k[5] = {1, 4, 6, 4, 1};
  int dw = (w + 1) / 2, dh = (h + 1) / 2;
  for (int y = 0; y < dh; y++)
    for (int x = 0; x < dw; x++) {
      int s = 0;
      for (int j = -2; j <= 2; j++) {
        const uint8_t* row = src + (size_t)refl101(2 * y + j, h) * w;
        int r = 0;
        for (int i = -2; i <= 2; i++) r += k[i + 2] * row[refl101(2 * x + i, w)];
        s += k[j + 2] * r;
      }
      dst[(size_t)y * dw + x] = (uint8_t)((s + 128) >> 8);
    }
}

typedef OrcKltLevel Level; /* vpl_oracle.h: w, h, pad, stride, img (REFLECT_101 border), deriv (zero border) */

static void level_from(const uint8_t* src, int w, int h, int pad, Level* L) {
  L->w = w; L->h = h; L->pad = pad; L->stride = w + 2 * pad;
  L->img = (uint8_t*)malloc((size_t)L->stride * (h + 2 * pad));
  L->deriv = NULL;
  for (int y = -pad; y < h + pad; y++) {
    const uint8_t* row = src + (size_t)refl101(y, h) * w;
    uint8_t* d = L->img + (size_t)(y + pad) * L->stride;
    for (int x = -pad; x < w + pad; x++) d[x + pad] = row[refl101(x, w)];
  }
}
/* KLT::calcSharrDeriv (klt.cpp:42-122) + copyMakeBorder(BORDER_CONSTANT) (klt.cpp:613) */
static void level_deriv(const uint8_t* src, Level* L) {
  int w = L->w, h = L->h, pad = L->pad;
  L->deriv = (int16_t*)calloc((size_t)L->stride * (h + 2 * pad) * 2, sizeof(int16_t));
  int* t0 = (int*)malloc(sizeof(int) * (w + 2));
  int* t1 = (int*)malloc(sizeof(int) * (w + 2));
  for (int y = 0; y < h; y++) {
    const uint8_t* r0 = src + (size_t)refl101(y - 1, h) * w;
    const uint8_t* r1 = src + (size_t)y * w;
    const uint8_t* r2 = src + (size_t)refl101(y + 1, h) * w;
    for (int x = 0; x < w; x++) {
      t0[x + 1] = (r0[x] + r2[x]) * 3 + r1[x] * 10;
      t1[x + 1] = r2[x] - r0[x];
    }
    t0[0] = t0[1 + refl101(-1, w)]; t0[w + 1] = t0[1 + refl101(w, w)];
    t1[0] = t1[1 + refl101(-1, w)]; t1[w + 1] = t1[1 + refl101(w, w)];
    int16_t* d = L->deriv + ((size_t)(y + pad) * L->stride + pad) * 2;
    for (int x = 0; x < w; x++) {
      d[2 * x] = (int16_t)(t0[x + 2] - t0[x]);
      d[2 * x + 1] = (int16_t)((t1[x + 2] + t1[x]) * 3 + t1[x + 1] * 10);
    }
  }
  free(t0); free(t1);
}

/* levels[0..return]; `with_deriv`: the reference computes the derivative of the prev pyramid only */
static int build_pyramid(const uint8_t* img, int w, int h, int win, int max_level, int with_deriv, Level* levels) {
  uint8_t* cur = (uint8_t*)malloc((size_t)w * h);
  memcpy(cur, img, (size_t)w * h);
  int cw = w, ch = h, level;
  for (level = 0; level <= max_level; level++) {
    level_from(cur, cw, ch, win, &levels[level]);
    if (with_deriv) level_deriv(cur, &levels[level]);
    int nw = (cw + 1) / 2, nh = (ch + 1) / 2;
    if (nw <= win || nh <= win) break; /* buildOpticalFlowPyramid returns this level */
    if (level == max_level) break;
    uint8_t* nxt = (uint8_t*)malloc((size_t)nw * nh);
    orc_pyrdown_std(cur, cw, ch, nxt);
    free(cur);
    cur = nxt; cw = nw; ch = nh;
  }
  free(cur);
  return level > max_level ? max_level : level;
}
static void free_levels(Level* L, int n) {
  for (int i = 0; i <= n; i++) { free(L[i].img); free(L[i].deriv); }
}
/* exported for oracle/ref_linematch_glue.cpp (the reference's tracker runs on these buffers) and
 * for the pyramid tests against cv2.buildOpticalFlowPyramid / cv2.Scharr */
int orc_klt_build_levels(const uint8_t* img, int w, int h, int win, int max_level, int with_deriv, OrcKltLevel* levels) {
  return build_pyramid(img, w, h, win, max_level > 7 ? 7 : max_level, with_deriv, levels);
}
void orc_klt_free_levels(OrcKltLevel* levels, int top) { free_levels(levels, top); }

/* ---- getImageNormParams (klt.cpp:4-10) on two win x win CV_16S windows ---------------------- */
static void mean_sd_16s(const int16_t* v, int n, double* mean, double* sd) { /* cv::meanStdDev */
  int s = 0;
  double q = 0;
  for (int i = 0; i < n; i++) { s += v[i]; q += (double)v[i] * v[i]; }
  double scale = 1.0 / n;
  *mean = s * scale;
  double var = q * scale - *mean * *mean;
  *sd = sqrt(var > 0 ? var : 0);
}
static void norm_params(const int16_t* I, const int16_t* J, int n, float* alpha, float* beta) {
  double sm, ss, dm, ds;
  mean_sd_16s(I, n, &sm, &ss);
  mean_sd_16s(J, n, &dm, &ds);
  *alpha = (float)(ss / ds);
  *beta = (float)(sm - *alpha * dm);
}

static void weights(float a, float b, int iw[4]) { /* lk2d.cpp:109-112 */
  iw[0] = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
  iw[1] = cv_round(a * (1.f - b) * (1 << W_BITS));
  iw[2] = cv_round((1.f - a) * b * (1 << W_BITS));
  iw[3] = (1 << W_BITS) - iw[0] - iw[1] - iw[2];
}
static void sample_window(const Level* L, int ix, int iy, const int iw[4], int win, int16_t* out) {
  /* lk2d.cpp:338-348: bilinear window of the (padded) image, 5 fractional bits kept */
  for (int y = 0; y < win; y++) {
    const uint8_t* p = L->img + (size_t)(y + iy + L->pad) * L->stride + (ix + L->pad);
    for (int x = 0; x < win; x++)
      out[y * win + x] = (int16_t)DESCALE(p[x] * iw[0] + p[x + 1] * iw[1] + p[x + L->stride] * iw[2] + p[x + L->stride + 1] * iw[3],
                                          W_BITS - 5);
  }
}

/* ---- LKTrackerInvoker2D::operator() for one point at one level (lk2d.cpp:28-480, affines == nullptr) */
typedef struct {
  int win, max_count, flags_unused;
  double epsilon; /* already squared, klt.cpp:33 */
  float min_eig;
  int illumination_adapt;
} LkParam;

static void lk_point(const Level* I, const Level* J, const LkParam* P, int level, int max_level, const float* prev_pt,
                     float* next_pt, uint8_t* status, float* err, int16_t* Iw, int16_t* dIw, int16_t* Jw) {
  const int win = P->win, n = win * win;
  const float half = (win - 1) * 0.5f;
  const float FLT_SCALE = 1.f / (1 << 20);
  float scale = (float)(1. / (1 << level));
  float px = prev_pt[0] * scale, py = prev_pt[1] * scale, nx, ny;
  if (level == max_level) { nx = px; ny = py; }            /* flags == 0: no initial flow, lk2d.cpp:51-56 */
  else { nx = next_pt[0] * 2.f; ny = next_pt[1] * 2.f; }   /* lk2d.cpp:58 */
  next_pt[0] = nx; next_pt[1] = ny;                        /* lk2d.cpp:61 */
  px -= half; py -= half;
  int ipx = cv_floor(px), ipy = cv_floor(py);
  if (ipx < -win || ipx >= I->w || ipy < -win || ipy >= I->h) { /* lk2d.cpp:91-99 */
    if (level == 0) { *status = 0; *err = 0; }
    return;
  }
  int iw[4];
  weights(px - ipx, py - ipy, iw);
  float iA11 = 0, iA12 = 0, iA22 = 0;
  for (int y = 0; y < win; y++) { /* lk2d.cpp:117-149 */
    const uint8_t* s = I->img + (size_t)(y + ipy + I->pad) * I->stride + (ipx + I->pad);
    const int16_t* d = I->deriv + ((size_t)(y + ipy + I->pad) * I->stride + (ipx + I->pad)) * 2;
    const int ds = I->stride * 2;
    for (int x = 0; x < win; x++, d += 2) {
      int ival = DESCALE(s[x] * iw[0] + s[x + 1] * iw[1] + s[x + I->stride] * iw[2] + s[x + I->stride + 1] * iw[3], W_BITS - 5);
      int ixval = DESCALE(d[0] * iw[0] + d[2] * iw[1] + d[ds] * iw[2] + d[ds + 2] * iw[3], W_BITS);
      int iyval = DESCALE(d[1] * iw[0] + d[3] * iw[1] + d[ds + 1] * iw[2] + d[ds + 3] * iw[3], W_BITS);
      Iw[y * win + x] = (int16_t)ival;
      dIw[2 * (y * win + x)] = (int16_t)ixval;
      dIw[2 * (y * win + x) + 1] = (int16_t)iyval;
      iA11 += (float)(ixval * ixval);
      iA12 += (float)(ixval * iyval);
      iA22 += (float)(iyval * iyval);
    }
  }
  float A11 = iA11 * FLT_SCALE, A12 = iA12 * FLT_SCALE, A22 = iA22 * FLT_SCALE;
  float D = A11 * A22 - A12 * A12;
  float min_eig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
  if (min_eig < P->min_eig || D < FLT_EPSILON) { /* lk2d.cpp:294-298 */
    if (level == 0) *status = 0;
    return;
  }
  D = 1.f / D;
  nx -= half; ny -= half;
  float pdx = 0, pdy = 0;
  int j;
  for (j = 0; j < P->max_count; j++) { /* lk2d.cpp:321-418 */
    int inx = cv_floor(nx), iny = cv_floor(ny);
    if (inx < -half || inx >= J->w || iny < -half || iny >= J->h) {
      if (level == 0) *status = 0;
      break;
    }
    weights(nx - inx, ny - iny, iw);
    sample_window(J, inx, iny, iw, win, Jw);
    float alpha = 1.0f, beta = 0.0f;
    if (P->illumination_adapt) norm_params(Iw, Jw, n, &alpha, &beta);
    float ib1 = 0, ib2 = 0;
    for (int i = 0; i < n; i++) { /* lk2d.cpp:364-376 */
      float diff = (float)(alpha * Jw[i] + beta - Iw[i]);
      ib1 += (float)(diff * dIw[2 * i]);
      ib2 += (float)(diff * dIw[2 * i + 1]);
    }
    float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
    float dx = (float)((A12 * b2 - A22 * b1) * D), dy = (float)((A12 * b1 - A11 * b2) * D);
    nx += dx; ny += dy;
    next_pt[0] = nx + half; next_pt[1] = ny + half;
    if ((double)dx * dx + (double)dy * dy <= P->epsilon) break; /* delta.ddot(delta), lk2d.cpp:405 */
    if (j > 0 && fabsf(dx + pdx) < 0.01 && fabsf(dy + pdy) < 0.01) { /* lk2d.cpp:410-414 */
      next_pt[0] -= dx * 0.5f; next_pt[1] -= dy * 0.5f;
      break;
    }
    pdx = dx; pdy = dy;
  }
  if (j == P->max_count && level == 0) *status = 0; /* lk2d.cpp:422 */
  if (level == 0 && *status) { /* final error, lk2d.cpp:429-478 */
    float ex = next_pt[0] - half, ey = next_pt[1] - half;
    int iex = cv_floor(ex), iey = cv_floor(ey);
    if (iex < -win || iex >= J->w || iey < -win || iey >= J->h) { *status = 0; return; }
    weights(ex - iex, ey - iey, iw);
    sample_window(J, iex, iey, iw, win, Jw);
    float alpha = 1.0f, beta = 0.0f;
    if (P->illumination_adapt) norm_params(Iw, Jw, n, &alpha, &beta);
    float errval = 0.f;
    for (int i = 0; i < n; i++) errval += fabsf((float)(alpha * Jw[i] + beta - Iw[i]));
    *err = errval * 1.f / (32 * win * win);
  }
}

/* ---- KLT::calc2D (klt.cpp:491-628), flags = 0 ----------------------------------------------- */
void orc_klt_calc2d(const uint8_t* img_ref, const uint8_t* img_cur, int w, int h, const float* prev_pts, int n,
                    int win, int max_level, int max_count, double epsilon, float min_eig, int illumination_adapt,
                    float* next_pts, uint8_t* status, float* err) {
  Level LI[8], LJ[8];
  if (max_level > 7) max_level = 7;
  int l1 = build_pyramid(img_ref, w, h, win, max_level, 1, LI);
  int l2 = build_pyramid(img_cur, w, h, win, l1, 0, LJ); /* same sizes => same depth */
  (void)l2;
  LkParam P = {win, max_count, 0, epsilon * epsilon, min_eig, illumination_adapt}; /* epsilon squared, klt.cpp:33 */
  int16_t* Iw = (int16_t*)malloc(sizeof(int16_t) * win * win * 4);
  int16_t *dIw = Iw + win * win, *Jw = Iw + 3 * win * win;
  for (int i = 0; i < n; i++) { status[i] = 1; err[i] = 0; next_pts[2 * i] = next_pts[2 * i + 1] = 0; }
  for (int level = l1; level >= 0; level--)
    for (int i = 0; i < n; i++)
      lk_point(&LI[level], &LJ[level], &P, level, l1, prev_pts + 2 * i, next_pts + 2 * i, status + i, err + i, Iw, dIw, Jw);
  free(Iw);
  free_levels(LI, l1);
  free_levels(LJ, l1);
}

/* ---- LineMatching (lm.cpp) -------------------------------------------------------------------- */
static float point_line_distance(float x, float y, const float e[4]) { /* lm.cpp:21-41 */
  float v_x = e[2] - e[0], v_y = e[3] - e[1];
  float u_x = e[0] - x, u_y = e[1] - y;
  float t = -(v_x * u_x + v_y * u_y) / (v_x * v_x + v_y * v_y);
  if (t < 0) t = 0;
  else if (t > 1) t = 1;
  float d_x = t * v_x + u_x, d_y = t * v_y + u_y;
  return sqrtf(d_x * d_x + d_y * d_y);
}

/* Anchors, lm.cpp:531-599.  Returns the number of anchor points. */
int orc_lm_anchors(const OrcLine* lines, int n_lines, int step, float* kps, int32_t* line_kp_num, int cap) {
  int n = 0;
  for (int i = 0; i < n_lines; i++) {
    float x1 = lines[i].endpoint[0], y1 = lines[i].endpoint[1], x2 = lines[i].endpoint[2], y2 = lines[i].endpoint[3];
    float px = x1, py = y1, len = lines[i].length;
    float dirx = (x2 - x1) / len, diry = (y2 - y1) / len;
    float ddx = step * dirx, ddy = step * diry;
    int iter = (int)(len / step);
    for (int j = 0; j <= iter; j++) {
      if (n < cap) { kps[2 * n] = px; kps[2 * n + 1] = py; }
      n++;
      px += ddx; py += ddy;
    }
    if (n < cap) { kps[2 * n] = x2; kps[2 * n + 1] = y2; }
    n++;
    line_kp_num[i] = iter + 2;
  }
  return n;
}

static int sideness_check(const OrcLine* l1r, const OrcLine* l2r, const OrcLine* l1c, const OrcLine* l2c, float* d1,
                          float* d2) { /* lm.cpp:412-446 */
  double a_1 = l1r->equation[0], b_1 = l1r->equation[1], c_1 = l1r->equation[2];
  double px_1 = l2r->center[0], py_1 = l2r->center[1];
  double a_2 = l1c->equation[0], b_2 = l1c->equation[1], c_2 = l1c->equation[2];
  double px_2 = l2c->center[0], py_2 = l2c->center[1];
  if ((fabs(a_1 - a_2) + fabs(b_1 - b_2)) > (fabs(a_1 + a_2) + fabs(b_1 + b_2))) { a_2 = -a_2; b_2 = -b_2; c_2 = -c_2; }
  *d1 = (float)((px_1 * a_1 + py_1 * b_1 + c_1) / sqrt(a_1 * a_1 + b_1 * b_1));
  *d2 = (float)((px_2 * a_2 + py_2 * b_2 + c_2) / sqrt(a_2 * a_2 + b_2 * b_2));
  return !(*d1 * *d2 < 0);
}

void orc_lm_default_param(OrcLineMatchParam* p) { /* lm.cpp:3-15, :630-631, line_matching.h:14-18, :45-47 */
  p->step = 10; p->closest_line_threshold = 0.5f; p->line_matching_ratio = 0.4f; p->line_distance_error_ratio = 3.f;
  p->klt_error_threshold = 40.f;
  p->win = 13; p->max_level = 3; p->max_count = 30; p->epsilon = 0.001; p->min_eig = 1e-4f;
  p->topo_distance_threshold = 15.f; p->topo_length_ratio = 0.2f; p->topo_violation_ratio = 0.05f;
}

/* LineMatching::Matching, lm.cpp:605-690.  Returns 0 when the reference returns false (an empty
 * input; ref_to_cur is then left untouched, as the reference leaves its vector), else 1. */
int orc_line_matching(const uint8_t* img_ref, const uint8_t* img_cur, int w, int h, const OrcLine* lines_ref, int n_ref,
                      const OrcLine* lines_cur, int n_cur, const OrcLineMatchParam* P, int illumination_adapt,
                      int topological_filter, int32_t* ref_to_cur, float* kps_ref_out, float* kps_cur_out,
                      uint8_t* status_out, float* err_out, int32_t* kp2line_out, int cap_kp, int* n_kp_out) {
  if (n_kp_out) *n_kp_out = 0;
  if (n_ref == 0 || n_cur == 0) return 0; /* lm.cpp:621 */
  int32_t* kp_num = (int32_t*)malloc(sizeof(int32_t) * n_ref);
  int n_kp = orc_lm_anchors(lines_ref, n_ref, P->step, NULL, kp_num, 0);
  float* kps = (float*)malloc(sizeof(float) * 2 * (n_kp + 1));
  orc_lm_anchors(lines_ref, n_ref, P->step, kps, kp_num, n_kp);
  float* nxt = (float*)malloc(sizeof(float) * 2 * (n_kp + 1));
  uint8_t* status = (uint8_t*)malloc(n_kp + 1);
  float* err = (float*)malloc(sizeof(float) * (n_kp + 1));
  int32_t* kp2line = (int32_t*)malloc(sizeof(int32_t) * (n_kp + 1));
  orc_klt_calc2d(img_ref, img_cur, w, h, kps, n_kp, P->win, P->max_level, P->max_count, P->epsilon, P->min_eig,
                 illumination_adapt, nxt, status, err);
  /* ClosestLine, lm.cpp:48-86 */
  for (int i = 0; i < n_kp; i++) {
    kp2line[i] = -1;
    if (!status[i] || err[i] > P->klt_error_threshold) continue;
    int min_idx = -1;
    float min_d = 1000000;
    for (int j = 0; j < n_cur; j++) {
      float d = point_line_distance(nxt[2 * i], nxt[2 * i + 1], lines_cur[j].endpoint);
      if (d < min_d) { min_d = d; min_idx = j; }
    }
    if (min_d < P->closest_line_threshold) kp2line[i] = min_idx;
  }
  /* Point2Line, lm.cpp:88-133 */
  int* count = (int*)malloc(sizeof(int) * n_cur);
  int idx = 0, max_idx = 0;
  for (int i = 0; i < n_ref; i++) {
    ref_to_cur[i] = -1;
    memset(count, 0, sizeof(int) * n_cur);
    for (int j = 0; j < kp_num[i]; j++, idx++)
      if (kp2line[idx] != -1) count[kp2line[idx]]++;
    int max_value = -1;
    for (int j = 0; j < n_cur; j++)
      if (count[j] > max_value) { max_value = count[j]; max_idx = j; }
    if (max_value <= 2 || (float)max_value / kp_num[i] < P->line_matching_ratio ||
        lines_cur[max_idx].length > lines_ref[i].length * P->line_distance_error_ratio ||
        lines_cur[max_idx].length < lines_ref[i].length / P->line_distance_error_ratio)
      continue;
    ref_to_cur[i] = max_idx;
  }
  free(count);
  if (topological_filter) { /* TopologicalFilter, lm.cpp:267-410, then lm.cpp:656-665 */
    int* viol = (int*)calloc(n_ref, sizeof(int));
    int match_num = 0;
    for (int r1 = 0; r1 < n_ref; r1++) {
      int c1 = ref_to_cur[r1];
      if (c1 == -1) continue;
      match_num++;
      for (int r2 = 0; r2 < n_ref; r2++) {
        if (r1 == r2) continue;
        int c2 = ref_to_cur[r2];
        if (c2 == -1) continue;
        float ratio = fabsf(lines_ref[r2].length - lines_cur[c2].length) / lines_ref[r2].length;
        if (ratio > P->topo_length_ratio) continue;
        float d_ref, d_cur;
        int same = sideness_check(&lines_ref[r1], &lines_ref[r2], &lines_cur[c1], &lines_cur[c2], &d_ref, &d_cur);
        if (!same && fabsf(d_ref) > P->topo_distance_threshold && fabsf(d_cur) > P->topo_distance_threshold) {
          viol[r1] += 1;
          viol[r2] += 1;
        }
      }
    }
    float threshold = P->topo_violation_ratio * (match_num - 1);
    if (threshold < 2) threshold = 2;
    for (int r = 0; r < n_ref; r++)
      if (viol[r] > threshold) ref_to_cur[r] = -1;
    free(viol);
  }
  if (n_kp_out) *n_kp_out = n_kp;
  for (int i = 0; i < n_kp && i < cap_kp; i++) {
    if (kps_ref_out) { kps_ref_out[2 * i] = kps[2 * i]; kps_ref_out[2 * i + 1] = kps[2 * i + 1]; }
    if (kps_cur_out) { kps_cur_out[2 * i] = nxt[2 * i]; kps_cur_out[2 * i + 1] = nxt[2 * i + 1]; }
    if (status_out) status_out[i] = status[i];
    if (err_out) err_out[i] = err[i];
    if (kp2line_out) kp2line_out[i] = kp2line[i];
  }
  free(kp_num); free(kps); free(nxt); free(status); free(err); free(kp2line);
  return 1;
}
