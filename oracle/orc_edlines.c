/*
 * orc_edlines.c -- CPU ORACLE (test infrastructure only, see vpl_oracle.h) for the reference's
 * real line detector, SURVEY.md 8(f)-1: EDLineDetector::EDline
 *   /root/reference/line_matching/src/edline_detector.cpp  (cited below as ed.cpp:line)
 *   /root/reference/line_matching/src/edline_detector.h    (ed.h:line)
 * called by the tracker at feature_tracker/src/line_feature_tracker.cpp:87 -> :315-321.
 *
 * A restatement, not a copy: one walk routine instead of the four unrolled ones, integer sums
 * instead of cv::Mat_ products (exactly equal, see fit_solve), explicit output order.
 * PINNED: tests/test_oracle_edlines.py compares every output (edge chains and lines, bit for bit)
 * with the reference's own edline_detector.cpp compiled against oracle/cvshim
 * (oracle/_ref/libref_linefront.so, built by `make -C oracle ref`) and with the golden vectors that
 * build produced (tests/golden/ref_edlines.npz).
 *
 * OpenCV calls inside EdgeDrawing (ed.cpp:125-136) are restated with the semantics probed on
 * cv2 4.13 (round-half-even `/ 4`, threshold on floor(thresh)); Sobel/GaussianBlur are
 * orc_prims.c (bit-exact vs cv2 4.13).
 *
 * Output order: the reference pushes lines under a mutex from parallel_for_ stripes
 * (ed.cpp:1081-1083, :1195) -- a race.  The order here is (edge chain, position in the chain),
 * i.e. what the reference gives on one thread.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "vpl_oracle.h"

#define ORC_PI 3.14159265358979323846 /* M_PI */
#define HORIZONTAL 255 /* |dx| < |dy|   ed.cpp:5   */
#define VERTICAL 0     /* |dy| <= |dx|  ed.cpp:6   */
enum { UP = 1, RIGHT = 2, DOWN = 3, LEFT = 4 }; /* ed.cpp:7-10 */
#define TRY_TIME 6       /* ed.cpp:11 */
#define SKIP_EDGE_POINT 2 /* ed.cpp:12 */

/* ---- gradient maps, ed.cpp:125-136 ------------------------------------------------------- */
static short div4_round_half_even(int v) { /* `Mat / 4` = convertTo(alpha .25f): cvRound */
  int q = v >> 2, r = v & 3;               /* v >= 0 here */
  if (r == 3 || (r == 2 && (q & 1))) q++;
  return (short)q;
}

void orc_ed_gradient_maps(const uint8_t* img, int w, int h, int smoothed, int grad_thresh,
                          int16_t* dx, int16_t* dy, int16_t* g, uint8_t* dir) {
  uint8_t* blurred = NULL;
  if (!smoothed) { /* ed.cpp:82-83 (ksize 5, sigma 1 only) */
    blurred = (uint8_t*)malloc((size_t)w * h);
    orc_gaussian_blur5(img, w, h, blurred);
    img = blurred;
  }
  orc_sobel3(img, w, h, dx, dy); /* ed.cpp:125-126 */
  for (size_t i = 0, n = (size_t)w * h; i < n; i++) {
    int ax = abs(dx[i]), ay = abs(dy[i]); /* ed.cpp:128-129 */
    int s = ax + ay;                      /* ed.cpp:131 */
    g[i] = div4_round_half_even(s > grad_thresh + 1 ? s : 0); /* ed.cpp:133-134 */
    dir[i] = ax < ay ? HORIZONTAL : VERTICAL;                 /* ed.cpp:136 */
  }
  free(blurred);
}

/* ---- one directional walk, ed.cpp:211-312 (and its three copies :320-421, :426-527, :535-636) */
typedef struct {
  int w, h;
  const int16_t* g;
  const uint8_t* dir;
  uint8_t* edge;
  unsigned last_x, last_y; /* carried across walks like the reference's locals (ed.cpp:184-185) */
} Walk;

static unsigned walk(Walk* s, unsigned x, unsigned y, int last_direction, uint32_t* out, unsigned n) {
  const int W = s->w, H = s->h;
  long idx = (long)y * W + x;
  while (s->g[idx] > 0 && !s->edge[idx]) {
    s->edge[idx] = 1;
    out[n++] = x | (y << 16);
    int should_go = 0;
    int g1, g2, g3;
    if (s->dir[idx] == HORIZONTAL) {
      if (last_direction == UP || last_direction == DOWN) should_go = x > s->last_x ? RIGHT : LEFT;
      s->last_x = x;
      s->last_y = y;
      if (last_direction == RIGHT || should_go == RIGHT) {
        if (x == (unsigned)W - 1 || y == 0 || y == (unsigned)H - 1) break;
        g1 = (unsigned char)s->g[idx - W + 1]; /* the (unsigned char) casts: ed.cpp:231-233 */
        g2 = (unsigned char)s->g[idx + 1];
        g3 = (unsigned char)s->g[idx + W + 1];
        if (g1 >= g2 && g1 >= g3) { x++; y--; }
        else if (g3 >= g2 && g3 >= g1) { x++; y++; }
        else x++;
        last_direction = RIGHT;
      } else if (last_direction == LEFT || should_go == LEFT) {
        if (x == 0 || y == 0 || y == (unsigned)H - 1) break;
        g1 = (unsigned char)s->g[idx - W - 1];
        g2 = (unsigned char)s->g[idx - 1];
        g3 = (unsigned char)s->g[idx + W - 1];
        if (g1 >= g2 && g1 >= g3) { x--; y--; }
        else if (g3 >= g2 && g3 >= g1) { x--; y++; }
        else x--;
        last_direction = LEFT;
      }
    } else {
      if (last_direction == RIGHT || last_direction == LEFT) should_go = y > s->last_y ? DOWN : UP;
      s->last_x = x;
      s->last_y = y;
      if (last_direction == DOWN || should_go == DOWN) {
        if (x == 0 || x == (unsigned)W - 1 || y == (unsigned)H - 1) break;
        g1 = (unsigned char)s->g[idx + W + 1];
        g2 = (unsigned char)s->g[idx + W];
        g3 = (unsigned char)s->g[idx + W - 1];
        if (g1 >= g2 && g1 >= g3) { x++; y++; }
        else if (g3 >= g2 && g3 >= g1) { x--; y++; }
        else y++;
        last_direction = DOWN;
      } else if (last_direction == UP || should_go == UP) {
        if (x == 0 || x == (unsigned)W - 1 || y == 0) break;
        g1 = (unsigned char)s->g[idx - W + 1];
        g2 = (unsigned char)s->g[idx - W];
        g3 = (unsigned char)s->g[idx - W - 1];
        if (g1 >= g2 && g1 >= g3) { x++; y--; }
        else if (g3 >= g2 && g3 >= g1) { x--; y--; }
        else y--;
        last_direction = UP;
      }
    }
    idx = (long)y * W + x;
  }
  return n;
}

/* ---- EdgeDrawing, ed.cpp:81-710.  chain_xy: x | y<<16 per pixel (capacity w*h/5 per part is
 * the reference's; this routine allocates 2x that so that the after-the-fact capacity checks of
 * ed.cpp:655-666 can be evaluated without the reference's out-of-bounds writes).
 * Returns 1, or -1 with *n_chains = *n_px = 0 when the reference returns -1. */
int orc_edge_drawing(const uint8_t* img, int w, int h, const OrcEDLineParam* p, int smoothed,
                     int16_t* dx, int16_t* dy, uint8_t* dir, uint32_t* chain_xy, uint32_t* chain_sid,
                     int* n_px, int* n_chains) {
  const unsigned pixel_num = (unsigned)w * h;
  const unsigned cap_px = pixel_num / 5, cap_edges = cap_px / 20; /* ed.cpp:92-93 */
  const short grad_thresh = (short)p->gradientThreshold;          /* member types, ed.h:118,122,126 */
  const unsigned char anchor_thresh = (unsigned char)p->anchorThreshold;
  const unsigned scan = (unsigned)p->scanIntervals;
  int16_t* g = (int16_t*)malloc(sizeof(int16_t) * pixel_num);
  uint8_t* edge = (uint8_t*)calloc(pixel_num, 1);
  *n_px = *n_chains = 0;
  orc_ed_gradient_maps(img, w, h, smoothed, grad_thresh, dx, dy, g, dir);

  /* anchors, column-major scan, ed.cpp:148-164 */
  uint32_t* anchors = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)(w / (scan ? scan : 1) + 1) * (h / (scan ? scan : 1) + 1) + 16));
  unsigned n_anchor = 0;
  for (unsigned x = 1; scan && x + 1 < (unsigned)w; x += scan)
    for (unsigned y = 1; y + 1 < (unsigned)h; y += scan) {
      long i = (long)y * w + x;
      int ok = dir[i] == HORIZONTAL ? (g[i] >= g[i - w] + anchor_thresh && g[i] >= g[i + w] + anchor_thresh)
                                    : (g[i] >= g[i - 1] + anchor_thresh && g[i] >= g[i + 1] + anchor_thresh);
      if (ok) anchors[n_anchor++] = x | (y << 16);
    }
  int status = 1;
  if (n_anchor > cap_px) status = -1; /* ed.cpp:166-169 */

  /* smart routing, ed.cpp:191-647 */
  /* every pixel is recorded at most once per mark, plus one re-walk of the anchor per chain */
  size_t room = (size_t)pixel_num + n_anchor + 16;
  uint32_t* first = (uint32_t*)malloc(sizeof(uint32_t) * room);
  uint32_t* second = (uint32_t*)malloc(sizeof(uint32_t) * room);
  unsigned* first_s = (unsigned*)malloc(sizeof(unsigned) * (n_anchor + 2));
  unsigned* second_s = (unsigned*)malloc(sizeof(unsigned) * (n_anchor + 2));
  unsigned off1 = 0, off2 = 0, n_edge = 0;
  Walk s = {w, h, g, dir, edge, 0, 0};
  for (unsigned a = 0; status == 1 && a < n_anchor; a++) {
    unsigned x = anchors[a] & 0xffff, y = anchors[a] >> 16;
    long idx = (long)y * w + x;
    if (edge[idx]) continue; /* ed.cpp:195 */
    first_s[n_edge] = off1;
    int horizontal = dir[idx] == HORIZONTAL;
    off1 = walk(&s, x, y, horizontal ? RIGHT : DOWN, first, off1);
    edge[idx] = 0; /* the anchor is walked again, ed.cpp:317 / :533 */
    second_s[n_edge] = off2;
    off2 = walk(&s, x, y, horizontal ? LEFT : UP, second, off2);
    if ((int)(off1 - first_s[n_edge]) + (int)(off2 - second_s[n_edge]) < p->minLineLen + 1) {
      off1 = first_s[n_edge]; /* short edge: records dropped, edge marks stay, ed.cpp:641-643 */
      off2 = second_s[n_edge];
    } else {
      n_edge++;
    }
  }
  first_s[n_edge] = off1;
  second_s[n_edge] = off2;
  if (n_edge > cap_edges) status = -1;                 /* ed.cpp:655 */
  if (off1 > cap_px || off2 > cap_px) status = -1;      /* ed.cpp:661 */
  if (!(off1 && off2)) status = -1;                     /* ed.cpp:667 */

  if (status == 1) { /* re-pack: reversed first part, then second part without the anchor, ed.cpp:687-706 */
    unsigned k = 0;
    for (unsigned e = 0; e < n_edge; e++) {
      chain_sid[e] = k;
      for (long t = (long)first_s[e + 1] - 1; t >= (long)first_s[e]; t--) chain_xy[k++] = first[t];
      for (unsigned t = second_s[e] + 1; t < second_s[e + 1]; t++) chain_xy[k++] = second[t];
    }
    chain_sid[n_edge] = k;
    *n_px = (int)k;
    *n_chains = (int)n_edge;
  }
  free(g); free(edge); free(anchors); free(first); free(second); free(first_s); free(second_s);
  return status;
}

/* ---- nfa, ed.h:171-348 (same formulas as LSD's) ---------------------------------------------- */
static int double_equal(double a, double b) { /* ed.h:171-187 */
  if (a == b) return 1;
  double abs_diff = fabs(a - b), aa = fabs(a), bb = fabs(b);
  double abs_max = aa > bb ? aa : bb;
  if (abs_max < DBL_MIN) abs_max = DBL_MIN;
  return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}
static double log_gamma_lanczos(double x) { /* ed.h:210-222 */
  static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705,
                              1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * log(x + 5.5) - (x + 5.5), b = 0.0;
  for (int n = 0; n < 7; n++) {
    a -= log(x + (double)n);
    b += q[n] * pow(x, (double)n);
  }
  return a + log(b);
}
static double log_gamma_windschitl(double x) { /* ed.h:238-240 */
  return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
}
static double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); } /* ed.h:44 */

double orc_ed_nfa(int n, int k, double p, double logNT) { /* ed.h:275-348 */
  const double tolerance = 0.1, ln10 = 2.30258509299404568402;
  if (n == 0 || k == 0) return -logNT;
  if (n == k) return -logNT - (double)n * log10(p);
  double p_term = p / (1.0 - p);
  double log1term = log_gamma((double)n + 1.0) - log_gamma((double)k + 1.0) - log_gamma((double)(n - k) + 1.0) +
                    (double)k * log(p) + (double)(n - k) * log(1.0 - p);
  double term = exp(log1term);
  if (double_equal(term, 0.0)) {
    if ((double)k > (double)n * p) return -log1term / ln10 - logNT;
    return -logNT;
  }
  double bin_tail = term;
  for (int i = k + 1; i <= n; i++) {
    double bin_term = (double)(n - i + 1) / (double)i;
    double mult_term = bin_term * p_term;
    term *= mult_term;
    bin_tail += term;
    if (bin_term < 1.0) {
      double err = term * ((1.0 - pow(mult_term, (double)(n - i + 1))) / (1.0 - mult_term) - 1.0);
      if (err < tolerance * fabs(-log10(bin_tail) - logNT) * bin_tail) break;
    }
  }
  return -log10(bin_tail) - logNT;
}

/* ---- least-squares fits, ed.cpp:729-891 ---------------------------------------------------- */
/* The reference forms A^T A and A^T v with cv::Mat_<float> products (cv::gemm on CV_32F: double
 * accumulators, one cast to float per element).  Every factor is an integer pixel coordinate, so
 * the double sums are exact integers; summing in int64 and casting once gives the same floats. */
typedef struct {
  float ata[4]; /* [sum u^2, sum u, sum u, n]   u = x (horizontal, y = a x + b) or y (vertical) */
  float atv[2]; /* [sum u v, sum v] */
} Normal;

static void normal_sums(const uint32_t* xy, unsigned s, unsigned e, int horizontal, float ata[4], float atv[2]) {
  int64_t suu = 0, su = 0, suv = 0, sv = 0;
  for (unsigned i = s; i < e; i++) {
    int64_t x = xy[i] & 0xffff, y = xy[i] >> 16;
    int64_t u = horizontal ? x : y, v = horizontal ? y : x;
    suu += u * u; su += u; suv += u * v; sv += v;
  }
  ata[0] = (float)(double)suu; ata[1] = (float)(double)su; ata[2] = (float)(double)su;
  ata[3] = (float)(double)(int64_t)(e - s);
  atv[0] = (float)(double)suv; atv[1] = (float)(double)sv;
}
static void fit_solve(const Normal* nm, double eq[2]) { /* ed.cpp:761-764 */
  const float* a = nm->ata;
  double coef = 1.0 / ((double)a[0] * (double)a[3] - (double)a[1] * (double)a[2]);
  eq[0] = coef * ((double)a[3] * (double)nm->atv[0] - (double)a[1] * (double)nm->atv[1]);
  eq[1] = coef * ((double)a[0] * (double)nm->atv[1] - (double)a[2] * (double)nm->atv[0]);
}
/* initial fit over min_len pixels, returns the fit error, ed.cpp:729-803 */
static double fit_initial(const uint32_t* xy, unsigned s, int min_len, int horizontal, Normal* nm, double eq[2]) {
  normal_sums(xy, s, s + (unsigned)min_len, horizontal, nm->ata, nm->atv); /* ed.cpp:757-758 */
  fit_solve(nm, eq);
  double err = 0;
  for (int i = 0; i < min_len; i++) { /* ed.cpp:767-770 / :796-799 */
    double x = (double)(xy[s + i] & 0xffff), y = (double)(xy[s + i] >> 16);
    double c = horizontal ? y - x * eq[0] - eq[1] : x - y * eq[0] - eq[1];
    err += c * c;
  }
  return sqrt(err);
}
/* incremental refit with the pixels [ns, e), ed.cpp:805-891 */
static void fit_update(const uint32_t* xy, unsigned ns, unsigned e, int horizontal, Normal* nm, double eq[2]) {
  if ((int)(e - ns) <= 0) return; /* ed.cpp:811-816: returns -1, equation untouched */
  float ata[4], atv[2];
  normal_sums(xy, ns, e, horizontal, ata, atv); /* ed.cpp:843-844 */
  for (int i = 0; i < 4; i++) nm->ata[i] = nm->ata[i] + ata[i]; /* float adds, ed.cpp:845-846 */
  for (int i = 0; i < 2; i++) nm->atv[i] = nm->atv[i] + atv[i];
  fit_solve(nm, eq);
}

/* ---- LineValidation, ed.cpp:893-958 ---------------------------------------------------------- */
static int line_validation(const uint32_t* xy, unsigned s, unsigned e, const int16_t* dx, const int16_t* dy,
                           int w, int h, const double eq[3], double logNT) {
  int n = (int)(e - s);
  int mgx = 0, mgy = 0;
  for (unsigned i = s; i < e; i++) {
    long idx = (long)(xy[i] >> 16) * w + (xy[i] & 0xffff);
    mgx += dx[idx];
    mgy += dy[idx];
  }
  double adx = fabs(eq[1]), ady = fabs(eq[0]);
  if (mgx == 0 && mgy == 0) return 0;
  float direction = 0.f; /* every (mgx,mgy) != (0,0) falls in exactly one quadrant below */
  if (mgx > 0 && mgy >= 0) direction = (float)atan2(-ady, adx);
  if (mgx <= 0 && mgy > 0) direction = (float)atan2(ady, adx);
  if (mgx < 0 && mgy <= 0) direction = (float)atan2(ady, -adx);
  if (mgx >= 0 && mgy < 0) direction = (float)atan2(-ady, -adx);
  if (fabs(direction) < 0.15 || ORC_PI - fabs(direction) < 0.15) /* ed.cpp:932-936 */
    if (fabs(eq[2]) < 10 || fabs((unsigned)h - fabs(eq[2])) < 10) return 0;
  if (fabs(fabs(direction) - ORC_PI * 0.5) < 0.15) /* ed.cpp:937-941 */
    if (fabs(eq[2]) < 10 || fabs((unsigned)w - fabs(eq[2])) < 10) return 0;
  int k = 0;
  for (unsigned i = s; i < e; i++) { /* ed.cpp:945-950 */
    long idx = (long)(xy[i] >> 16) * w + (xy[i] & 0xffff);
    double pd = atan2(-(double)dx[idx], (double)dy[idx]); /* ed.cpp:912 */
    double dis = fabs(direction - pd);
    if (fabs(2 * ORC_PI - dis) < 0.392699 || dis < 0.392699) k++;
  }
  return orc_ed_nfa(n, k, 0.125, logNT) > 0; /* ed.cpp:952-954 */
}

static void emit_line(const uint32_t* xy, unsigned s0, unsigned s, const double eq[3], OrcLine* out) {
  /* ed.cpp:1056-1079 (and :1140-1162) */
  double a1 = eq[1] * eq[1], a2 = eq[0] * eq[0], a3 = eq[0] * eq[1], a4 = eq[2] * eq[0], a5 = eq[2] * eq[1];
  unsigned px = xy[s0] & 0xffff, py = xy[s0] >> 16;
  float x1 = (float)(a1 * px - a3 * py - a4), y1 = (float)(a2 * py - a3 * px - a5);
  px = xy[s - 1] & 0xffff; py = xy[s - 1] >> 16;
  float x2 = (float)(a1 * px - a3 * py - a4), y2 = (float)(a2 * py - a3 * px - a5);
  out->endpoint[0] = x1; out->endpoint[1] = y1; out->endpoint[2] = x2; out->endpoint[3] = y2;
  out->equation[0] = eq[0]; out->equation[1] = eq[1]; out->equation[2] = eq[2];
  out->center[0] = (float)((x1 + x2) / 2.0);
  out->center[1] = (float)((y1 + y2) / 2.0);
  double ddx = (double)(x2 - x1), ddy = (double)(y2 - y1); /* pow(float, 2) -> exact square in double */
  out->length = (float)sqrt(ddx * ddx + ddy * ddy);
  out->pad_ = 0;
}

/* ---- per-chain line extraction, ed.cpp:983-1173 -------------------------------------------- */
static int chain_lines(const uint32_t* xy, unsigned S, unsigned E, const uint8_t* dir, const int16_t* dx,
                       const int16_t* dy, int w, int h, const OrcEDLineParam* p, double logNT, OrcLine* out,
                       int n_out, int cap) {
  const int min_len = p->minLineLen;
  const double thr = p->lineFitErrThreshold;
  double fit_err = 0, eq[2] = {0, 0};
  Normal nm;
  while (E > S + (unsigned)min_len) { /* ed.cpp:987 */
    int horizontal = 0;
    while (E > S + (unsigned)min_len) { /* ed.cpp:989-995 */
      horizontal = dir[(long)(xy[S] >> 16) * w + (xy[S] & 0xffff)] == HORIZONTAL;
      fit_err = fit_initial(xy, S, min_len, horizontal, &nm, eq);
      if (fit_err <= thr) break;
      S += SKIP_EDGE_POINT;
    }
    if (fit_err > thr) break; /* ed.cpp:996 */
    /* after a skip past the end test the direction is re-read at the new S, ed.cpp:1005 */
    horizontal = dir[(long)(xy[S] >> 16) * w + (xy[S] & 0xffff)] == HORIZONTAL;
    double coef1 = 0;
    int extended = 1, first_try = 1, tries = 0, outliers;
    unsigned S0 = S, new_s = 0;
    while (extended) { /* ed.cpp:1008-1039 / :1090-1120 */
      tries++;
      if (first_try) {
        first_try = 0;
        S += (unsigned)min_len;
      } else {
        fit_update(xy, new_s, S, horizontal, &nm, eq);
      }
      coef1 = 1 / sqrt(horizontal ? eq[0] * eq[0] + 1 : 1 + eq[0] * eq[0]);
      outliers = 0;
      new_s = S;
      while (E > S) {
        unsigned X = xy[S] & 0xffff, Y = xy[S] >> 16;
        double d = horizontal ? fabs(eq[0] * X - Y + eq[1]) * coef1 : fabs(X - eq[0] * Y - eq[1]) * coef1;
        S++;
        if (d > thr) {
          if (++outliers > 3) break;
        } else {
          outliers = 0;
        }
      }
      S -= (unsigned)outliers;
      if (!((int)(S - new_s) > 0 && tries < TRY_TIME)) extended = 0; /* unsigned `> 0`, ed.cpp:1035 */
    }
    double le[3];
    if (horizontal) { /* ed.cpp:1041-1044 */
      le[0] = eq[0] * coef1; le[1] = -1 * coef1; le[2] = eq[1] * coef1;
    } else { /* ed.cpp:1122-1125 */
      le[0] = 1 * coef1; le[1] = -eq[0] * coef1; le[2] = -eq[1] * coef1;
    }
    if (line_validation(xy, S0, S, dx, dy, w, h, le, logNT)) {
      if (n_out < cap) emit_line(xy, S0, S, le, &out[n_out]);
      n_out++;
    }
  }
  return n_out;
}

/* ---- EDLineDetector::EDline, ed.cpp:1176-1198 ---------------------------------------------- */
int orc_edline_detect(const uint8_t* img, int w, int h, const OrcEDLineParam* p, int smoothed, OrcLine* out,
                      int cap, uint32_t* chain_xy_out, uint32_t* chain_sid_out, int* n_px_out, int* n_chains_out) {
  size_t n = (size_t)w * h;
  int16_t* dx = (int16_t*)malloc(2 * n);
  int16_t* dy = (int16_t*)malloc(2 * n);
  uint8_t* dir = (uint8_t*)malloc(n);
  uint32_t* xy = (uint32_t*)malloc(sizeof(uint32_t) * (n / 5 * 2 + 16));
  uint32_t* sid = (uint32_t*)calloc(n / 100 + 16, sizeof(uint32_t));
  int n_px = 0, n_chains = 0, n_lines = 0;
  int st = orc_edge_drawing(img, w, h, p, smoothed, dx, dy, dir, xy, sid, &n_px, &n_chains);
  /* on -1 the reference carries on with whatever edges_ held (ed.cpp:1180 never fires); with an
   * empty edges_ that is "no lines", which is what the oracle defines. */
  if (st == 1) {
    double logNT = 2.0 * (log10((double)(unsigned)w) + log10((double)(unsigned)h)); /* ed.cpp:1186 */
    for (int e = 0; e < n_chains; e++)
      n_lines = chain_lines(xy, sid[e], sid[e + 1], dir, dx, dy, w, h, p, logNT, out, n_lines, cap);
  }
  if (n_px_out) *n_px_out = n_px;
  if (n_chains_out) *n_chains_out = n_chains;
  if (chain_xy_out) memcpy(chain_xy_out, xy, sizeof(uint32_t) * (size_t)n_px);
  if (chain_sid_out) memcpy(chain_sid_out, sid, sizeof(uint32_t) * ((size_t)n_chains + 1));
  free(dx); free(dy); free(dir); free(xy); free(sid);
  return n_lines;
}
