/*
 * orc_lsd.c -- CPU oracle, cv::LineSegmentDetector restated in C.  TEST
 * INFRASTRUCTURE ONLY (see vpl_oracle.h).
 *
 * Follows opencv imgproc lsd.cpp (third party, un-vendored; the reference pins
 * "OpenCV 3.4.2", /root/reference/README.md:5) with the behaviour of cv2 4.13 --
 * the only obtainable build -- as arbiter (rules A.1-A.9).
 * log_gamma / nfa are the formulas also found in the reference at
 * line_matching/src/edline_detector.h:210-240 and :275-348.
 * Pinned bit-for-bit (segment count, order, endpoints, width, nfa) against
 * cv2.createLineSegmentDetector by tests/test_oracle_lsd.py and tests/golden/.
 * Build with -ffp-contract=off: the float32 polynomial of fastAtan2 and the
 * double sums must not be fused.
 */
#include "vpl_oracle.h"
#include "orc_sincos.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NOTDEF (-1024.0)
#define M_3_2_PI 4.71238898038468985769 /* (3 * CV_PI) / 2 */
#define M_2__PI 6.28318530717958647692  /* 2 * CV_PI */
#define ORC_PI 3.1415926535897932384626433832795
#define DEG_TO_RADS (ORC_PI / 180)
#define RELATIVE_ERROR_FACTOR 100.0
#define LN10 2.30258509299404568402

typedef struct {
  int x, y;
  double angle, modgrad;
} RegPt;

typedef struct {
  double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p;
} Rect;

typedef struct {
  int w, h; /* scaled image size */
  double* angles;
  double* modgrad;
  uint8_t* used;
  int* order; /* pixel indices y*w+x in processing order */
  int n_order;
  double log_nt;
  RegPt* reg;
  int nreg;
} Lsd;

static inline double dist(double x1, double y1, double x2, double y2) {
  return sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
}
static inline double dist_sq(double x1, double y1, double x2, double y2) {
  return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1);
}
static inline double angle_diff_signed(double a, double b) {
  double d = a - b;
  while (d <= -ORC_PI) d += M_2__PI;
  while (d > ORC_PI) d -= M_2__PI;
  return d;
}
static inline double angle_diff(double a, double b) { return fabs(angle_diff_signed(a, b)); }

static inline int double_equal(double a, double b) {
  if (a == b) return 1;
  double abs_diff = fabs(a - b);
  double aa = fabs(a), bb = fabs(b);
  double abs_max = aa > bb ? aa : bb;
  if (abs_max < DBL_MIN) abs_max = DBL_MIN;
  return (abs_diff / abs_max) <= (RELATIVE_ERROR_FACTOR * DBL_EPSILON);
}

/* edline_detector.h:210-240 carry the same two formulas */
static double log_gamma_lanczos(double x) {
  static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705,
                              1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
  double b = 0;
  for (int n = 0; n < 7; ++n) {
    a -= log(x + (double)n);
    b += q[n] * pow(x, (double)n);
  }
  return a + log(b);
}
static double log_gamma_windschitl(double x) {
  return 0.918938533204673 + (x - 0.5) * log(x) - x +
         0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
}
static double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); }

/* edline_detector.h:275-348 */
static double nfa(const Lsd* L, int n, int k, double p) {
  const double LOG_NT = L->log_nt;
  if (n == 0 || k == 0) return -LOG_NT;
  if (n == k) return -LOG_NT - (double)n * log10(p);
  double p_term = p / (1 - p);
  double log1term = log_gamma((double)n + 1) - log_gamma((double)k + 1) -
                    log_gamma((double)(n - k) + 1) + (double)k * log(p) +
                    (double)(n - k) * log(1.0 - p);
  double term = exp(log1term);
  if (double_equal(term, 0)) {
    if (k > n * p) return -log1term / LN10 - LOG_NT;
    else return -LOG_NT;
  }
  double bin_tail = term;
  double tolerance = 0.1;
  for (int i = k + 1; i <= n; ++i) {
    double bin_term = (double)(n - i + 1) / (double)i;
    double mult_term = bin_term * p_term;
    term *= mult_term;
    bin_tail += term;
    if (bin_term < 1) {
      double err = term * ((1 - pow(mult_term, (double)(n - i + 1))) / (1 - mult_term) - 1);
      if (err < tolerance * fabs(-log10(bin_tail) - LOG_NT) * bin_tail) break;
    }
  }
  return -log10(bin_tail) - LOG_NT;
}

static inline int is_aligned(const Lsd* L, int x, int y, double theta, double prec) {
  if (x < 0 || y < 0 || x >= L->w || y >= L->h) return 0;
  double a = L->angles[(size_t)y * L->w + x];
  if (a == NOTDEF) return 0;
  double n_theta = theta - a;
  if (n_theta < 0) n_theta = -n_theta;
  if (n_theta > M_3_2_PI) {
    n_theta -= M_2__PI;
    if (n_theta < 0) n_theta = -n_theta;
  }
  return n_theta <= prec;
}

/* A.3: gradient, level-line angle, pseudo-ordering (descending bin, raster
 * order inside a bin -- what cv2 4.13 produces, verified). */
static void ll_angle(Lsd* L, const uint8_t* img, double threshold, int n_bins) {
  const int w = L->w, h = L->h;
  for (int x = 0; x < w; ++x) { L->angles[(size_t)(h - 1) * w + x] = NOTDEF; L->modgrad[(size_t)(h - 1) * w + x] = 0; }
  for (int y = 0; y < h; ++y) { L->angles[(size_t)y * w + w - 1] = NOTDEF; L->modgrad[(size_t)y * w + w - 1] = 0; }
  double max_grad = -1;
  for (int y = 0; y < h - 1; ++y) {
    const uint8_t* r0 = img + (size_t)y * w;
    const uint8_t* r1 = img + (size_t)(y + 1) * w;
    for (int x = 0; x < w - 1; ++x) {
      int DA = r1[x + 1] - r0[x];
      int BC = r0[x + 1] - r1[x];
      int gx = DA + BC, gy = DA - BC;
      double norm = sqrt((gx * gx + gy * gy) / 4.0);
      L->modgrad[(size_t)y * w + x] = norm;
      if (norm <= threshold) L->angles[(size_t)y * w + x] = NOTDEF;
      else {
        L->angles[(size_t)y * w + x] = orc_fast_atan2((float)gx, (float)-gy) * DEG_TO_RADS;
        if (norm > max_grad) max_grad = norm;
      }
    }
  }
  double bin_coef = (max_grad > 0) ? (double)(n_bins - 1) / max_grad : 0;
  int* count = (int*)calloc((size_t)n_bins + 1, sizeof(int));
  int* bins = (int*)malloc((size_t)w * h * sizeof(int));
  for (int y = 0; y < h - 1; ++y)
    for (int x = 0; x < w - 1; ++x) {
      int b = (int)(L->modgrad[(size_t)y * w + x] * bin_coef);
      bins[(size_t)y * w + x] = b;
      count[b]++;
    }
  /* descending bins: start[b] = number of points with bin > b */
  int* start = (int*)malloc((size_t)n_bins * sizeof(int));
  int acc = 0;
  for (int b = n_bins - 1; b >= 0; --b) { start[b] = acc; acc += count[b]; }
  L->n_order = acc;
  for (int y = 0; y < h - 1; ++y)
    for (int x = 0; x < w - 1; ++x) {
      int b = bins[(size_t)y * w + x];
      L->order[start[b]++] = y * w + x;
    }
  free(count); free(bins); free(start);
}

/* A.4 */
static void region_grow(Lsd* L, int sx, int sy, double* reg_angle_, double prec) {
  const int w = L->w, h = L->h;
  RegPt* reg = L->reg;
  int n = 0;
  double reg_angle = L->angles[(size_t)sy * w + sx];
  reg[n].x = sx; reg[n].y = sy; reg[n].angle = reg_angle; reg[n].modgrad = L->modgrad[(size_t)sy * w + sx];
  n++;
  float sumdx = (float)cos(reg_angle);
  float sumdy = (float)sin(reg_angle);
  L->used[(size_t)sy * w + sx] = 1;
  for (int i = 0; i < n; ++i) {
    int px = reg[i].x, py = reg[i].y;
    int xx_min = px - 1 > 0 ? px - 1 : 0, xx_max = px + 1 < w - 1 ? px + 1 : w - 1;
    int yy_min = py - 1 > 0 ? py - 1 : 0, yy_max = py + 1 < h - 1 ? py + 1 : h - 1;
    for (int yy = yy_min; yy <= yy_max; ++yy)
      for (int xx = xx_min; xx <= xx_max; ++xx) {
        uint8_t* u = &L->used[(size_t)yy * w + xx];
        if (*u != 1 && is_aligned(L, xx, yy, reg_angle, prec)) {
          double angle = L->angles[(size_t)yy * w + xx];
          *u = 1;
          reg[n].x = xx; reg[n].y = yy; reg[n].angle = angle; reg[n].modgrad = L->modgrad[(size_t)yy * w + xx];
          n++;
          /* cos(float)/sin(float): the float overloads.  Restated as the
           * correctly rounded float result (float)cos((double)a), which is
           * what both glibc cosf (< 1 ulp) and the device path compute. */
          sumdx += (float)cos((double)(float)angle);
          sumdy += (float)sin((double)(float)angle);
          reg_angle = orc_fast_atan2(sumdy, sumdx) * DEG_TO_RADS;
        }
      }
  }
  L->nreg = n;
  *reg_angle_ = reg_angle;
}

/* A.5 */
static double get_theta(const Lsd* L, double x, double y, double reg_angle, double prec) {
  const RegPt* reg = L->reg;
  double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
  for (int i = 0; i < L->nreg; ++i) {
    double regx = reg[i].x, regy = reg[i].y, weight = reg[i].modgrad;
    double dx = regx - x, dy = regy - y;
    Ixx += dy * dy * weight;
    Iyy += dx * dx * weight;
    Ixy -= dx * dy * weight;
  }
  double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)orc_fast_atan2((float)(lambda - Ixx), (float)Ixy)
                                         : (double)orc_fast_atan2((float)Ixy, (float)(lambda - Iyy));
  theta *= DEG_TO_RADS;
  if (angle_diff(theta, reg_angle) > prec) theta += ORC_PI;
  return theta;
}

static void region2rect(const Lsd* L, double reg_angle, double prec, double p, Rect* rec) {
  const RegPt* reg = L->reg;
  double x = 0, y = 0, sum = 0;
  for (int i = 0; i < L->nreg; ++i) {
    double weight = reg[i].modgrad;
    x += (double)reg[i].x * weight;
    y += (double)reg[i].y * weight;
    sum += weight;
  }
  x /= sum;
  y /= sum;
  double theta = get_theta(L, x, y, reg_angle, prec);
  /* cos/sin through the deterministic correctly-rounded sincos (orc_sincos.h): the
   * rectangle edges pass through pixel centres, so rect_nfa's ceil()/trunc() limits are
   * sensitive to the last bit of dx, dy; libm is not correctly rounded in ~0.1 % of
   * cases and the device library in more.  ORC_LIBM_TRIG=1 restores plain libm. */
  double dx, dy;
#ifdef ORC_LIBM_TRIG
  dx = cos(theta); dy = sin(theta);
#else
  vpl_sincos_cr(theta, &dy, &dx);
#endif
  double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
  for (int i = 0; i < L->nreg; ++i) {
    double regdx = (double)reg[i].x - x;
    double regdy = (double)reg[i].y - y;
    double l = regdx * dx + regdy * dy;
    double w = regdy * dx - regdx * dy;
    if (l > l_max) l_max = l;
    else if (l < l_min) l_min = l;
    if (w > w_max) w_max = w;
    else if (w < w_min) w_min = w;
  }
  rec->x1 = x + l_min * dx;
  rec->y1 = y + l_min * dy;
  rec->x2 = x + l_max * dx;
  rec->y2 = y + l_max * dy;
  rec->width = w_max - w_min;
  rec->x = x; rec->y = y; rec->theta = theta; rec->dx = dx; rec->dy = dy;
  rec->prec = prec; rec->p = p;
  if (rec->width < 1.0) rec->width = 1.0;
}

/* A.6 */
static int reduce_region_radius(Lsd* L, double reg_angle, double prec, double p, Rect* rec,
                                double density, double density_th) {
  RegPt* reg = L->reg;
  const int w = L->w;
  double xc = (double)reg[0].x, yc = (double)reg[0].y;
  double radSq1 = dist_sq(xc, yc, rec->x1, rec->y1);
  double radSq2 = dist_sq(xc, yc, rec->x2, rec->y2);
  double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
  while (density < density_th) {
    radSq *= 0.75 * 0.75;
    for (int i = 0; i < L->nreg; ++i) {
      if (dist_sq(xc, yc, (double)reg[i].x, (double)reg[i].y) > radSq) {
        L->used[(size_t)reg[i].y * w + reg[i].x] = 0;
        RegPt t = reg[i]; reg[i] = reg[L->nreg - 1]; reg[L->nreg - 1] = t;
        L->nreg--;
        --i;
      }
    }
    if (L->nreg < 2) return 0;
    region2rect(L, reg_angle, prec, p, rec);
    density = (double)L->nreg / (dist(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
  }
  return 1;
}

static int refine(Lsd* L, double reg_angle, double prec, double p, Rect* rec, double density_th) {
  RegPt* reg = L->reg;
  const int w = L->w;
  double density = (double)L->nreg / (dist(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
  if (density >= density_th) return 1;
  double xc = (double)reg[0].x, yc = (double)reg[0].y;
  double ang_c = reg[0].angle;
  double sum = 0, s_sum = 0;
  int n = 0;
  for (int i = 0; i < L->nreg; ++i) {
    L->used[(size_t)reg[i].y * w + reg[i].x] = 0;
    if (dist(xc, yc, (double)reg[i].x, (double)reg[i].y) < rec->width) {
      double ang_d = angle_diff_signed(reg[i].angle, ang_c);
      sum += ang_d;
      s_sum += ang_d * ang_d;
      ++n;
    }
  }
  double mean_angle = sum / (double)n;
  double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)n + mean_angle * mean_angle);
  region_grow(L, reg[0].x, reg[0].y, &reg_angle, tau);
  if (L->nreg < 2) return 0;
  region2rect(L, reg_angle, prec, p, rec);
  density = (double)L->nreg / (dist(rec->x1, rec->y1, rec->x2, rec->y2) * rec->width);
  if (density < density_th) return reduce_region_radius(L, reg_angle, prec, p, rec, density, density_th);
  return 1;
}

/* A.7: cv2 4.13 rect_nfa (double vertices, ceil-guarded slopes). */
typedef struct { double x, y; } Pt;
static inline double slope_g(Pt a, Pt b) {
  return ((int)ceil(a.y) != (int)ceil(b.y)) ? (b.x - a.x) / (b.y - a.y) : 0;
}
static double rect_nfa(const Lsd* L, const Rect* rec) {
  int total_pts = 0, alg_pts = 0;
  double half_width = rec->width / 2.0;
  double dyhw = rec->dy * half_width;
  double dxhw = rec->dx * half_width;
  Pt v[4];
  v[0].x = rec->x1 - dyhw; v[0].y = rec->y1 + dxhw;
  v[1].x = rec->x2 - dyhw; v[1].y = rec->y2 + dxhw;
  v[2].x = rec->x2 + dyhw; v[2].y = rec->y2 - dxhw;
  v[3].x = rec->x1 + dyhw; v[3].y = rec->y1 - dxhw;
  /* rotate so that o[0] has the smallest y (ties: smallest x) */
  int offset = 0;
  {
    /* cv2: choose the offset by comparing coordinates of the four vertices */
    int best = 0;
    for (int i = 1; i < 4; ++i) {
      if (v[i].y < v[best].y || (v[i].y == v[best].y && v[i].x < v[best].x)) best = i;
    }
    offset = best;
  }
  Pt o[4];
  for (int i = 0; i < 4; ++i) o[i] = v[(offset + i) % 4];
  double flstep = slope_g(o[0], o[1]);
  double slstep = slope_g(o[1], o[2]);
  double frstep = slope_g(o[0], o[3]);
  double srstep = slope_g(o[3], o[2]);
  double top_y = o[0].y, bottom_y = o[2].y;
  int y1c = (int)ceil(o[1].y), y3c = (int)ceil(o[3].y);
  for (int y = (int)ceil(top_y); y <= (int)ceil(bottom_y); ++y) {
    if (y < 0 || y >= L->h) continue;
    double left = (y <= y1c) ? o[0].x + ((double)y - o[0].y) * flstep : o[1].x + ((double)y - o[1].y) * slstep;
    double right = (y < y3c) ? o[0].x + ((double)y - o[0].y) * frstep : o[3].x + ((double)y - o[3].y) * srstep;
    int xs = (int)ceil(left), xe = (int)right;
    /* same pixels as "for x = xs..xe: if (x < 0 || x >= w) continue", without walking the part of a
     * near-horizontal edge's span that lies outside the image (it can be 2^31 columns long) */
    if (xs < 0) xs = 0;
    if (xe > L->w - 1) xe = L->w - 1;
    for (int x = xs; x <= xe; ++x) {
      ++total_pts;
      if (is_aligned(L, x, y, rec->theta, rec->prec)) ++alg_pts;
    }
  }
  return nfa(L, total_pts, alg_pts, rec->p);
}

static double rect_improve(const Lsd* L, Rect* rec) {
  const double LOG_EPS = 0;
  double delta = 0.5, delta_2 = delta / 2.0;
  double log_nfa = rect_nfa(L, rec);
  if (log_nfa > LOG_EPS) return log_nfa;
  Rect r = *rec;
  for (int n = 0; n < 5; ++n) {
    r.p /= 2;
    r.prec = r.p * ORC_PI;
    double v = rect_nfa(L, &r);
    if (v > log_nfa) { log_nfa = v; *rec = r; }
  }
  if (log_nfa > LOG_EPS) return log_nfa;
  r = *rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.width -= delta;
      double v = rect_nfa(L, &r);
      if (v > log_nfa) { *rec = r; log_nfa = v; }
    }
  }
  if (log_nfa > LOG_EPS) return log_nfa;
  r = *rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2;
      r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
      r.width -= delta;
      double v = rect_nfa(L, &r);
      if (v > log_nfa) { *rec = r; log_nfa = v; }
    }
  }
  if (log_nfa > LOG_EPS) return log_nfa;
  r = *rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2;
      r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
      r.width -= delta;
      double v = rect_nfa(L, &r);
      if (v > log_nfa) { *rec = r; log_nfa = v; }
    }
  }
  if (log_nfa > LOG_EPS) return log_nfa;
  r = *rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.p /= 2;
      r.prec = r.p * ORC_PI;
      double v = rect_nfa(L, &r);
      if (v > log_nfa) { *rec = r; log_nfa = v; }
    }
  }
  return log_nfa;
}

/* debug sink: candidate rectangles after refine(), before rect_improve():
 * 16 doubles each = x1 y1 x2 y2 width x y theta dx dy prec p | nfa accepted region_size seed_index */
static double* g_cand_sink = NULL;
static int g_cand_cap = 0, g_cand_n = 0;

int orc_lsd_detect(const uint8_t* img, int w0, int h0, int refine_mode, int scale08, float* seg4,
                   double* width_out, double* prec_out, double* nfa_out, int cap) {
  const double ANG_TH = 22.5, QUANT = 2.0, DENSITY_TH = 0.7, LOG_EPS = 0, SCALE = 0.8;
  const int N_BINS = 1024;
  const double prec = ORC_PI * ANG_TH / 180;
  const double p = ANG_TH / 180;
  const double rho = QUANT / sin(prec);

  Lsd L;
  memset(&L, 0, sizeof(L));
  uint8_t* scaled;
  if (scale08) {
    uint8_t* g = (uint8_t*)malloc((size_t)w0 * h0);
    orc_gaussian_blur7_s075(img, w0, h0, g);
    orc_resize_08(g, w0, h0, NULL, &L.w, &L.h);
    scaled = (uint8_t*)malloc((size_t)L.w * L.h);
    orc_resize_08(g, w0, h0, scaled, &L.w, &L.h);
    free(g);
  } else {
    L.w = w0; L.h = h0;
    scaled = (uint8_t*)malloc((size_t)w0 * h0);
    memcpy(scaled, img, (size_t)w0 * h0);
  }
  const size_t np = (size_t)L.w * L.h;
  L.angles = (double*)malloc(np * sizeof(double));
  L.modgrad = (double*)malloc(np * sizeof(double));
  L.used = (uint8_t*)calloc(np, 1);
  L.order = (int*)malloc(np * sizeof(int));
  L.reg = (RegPt*)malloc(np * sizeof(RegPt));
  ll_angle(&L, scaled, rho, N_BINS);
  L.log_nt = 5 * (log10((double)L.w) + log10((double)L.h)) / 2 + log10(11.0);
  const int min_reg_size = (int)(-L.log_nt / log10(p));

  int nout = 0;
  for (int i = 0; i < L.n_order; ++i) {
    int pi = L.order[i];
    if (L.used[pi] != 0 || L.angles[pi] == NOTDEF) continue;
    double reg_angle;
    region_grow(&L, pi % L.w, pi / L.w, &reg_angle, prec);
    if (L.nreg < min_reg_size) continue;
    Rect rec;
    region2rect(&L, reg_angle, prec, p, &rec);
    double log_nfa = -1;
    if (refine_mode > 0) {
      if (!refine(&L, reg_angle, prec, p, &rec, DENSITY_TH)) continue;
      if (refine_mode >= 2) {
        double* sink = NULL;
        if (g_cand_sink && g_cand_n < g_cand_cap) {
          sink = g_cand_sink + 16 * (size_t)g_cand_n;
          const double v[12] = {rec.x1, rec.y1, rec.x2, rec.y2, rec.width, rec.x, rec.y, rec.theta, rec.dx, rec.dy, rec.prec, rec.p};
          memcpy(sink, v, sizeof(v));
          sink[14] = (double)L.nreg; sink[15] = (double)pi;
        }
        g_cand_n++;
        log_nfa = rect_improve(&L, &rec);
        if (sink) { sink[12] = log_nfa; sink[13] = (log_nfa > LOG_EPS) ? 1.0 : 0.0; }
        if (log_nfa <= LOG_EPS) continue;
      }
    }
    rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
    if (scale08) {
      rec.x1 /= SCALE; rec.y1 /= SCALE; rec.x2 /= SCALE; rec.y2 /= SCALE; rec.width /= SCALE;
    }
    if (nout < cap) {
      seg4[4 * nout + 0] = (float)rec.x1; seg4[4 * nout + 1] = (float)rec.y1;
      seg4[4 * nout + 2] = (float)rec.x2; seg4[4 * nout + 3] = (float)rec.y2;
      if (width_out) width_out[nout] = rec.width;
      if (prec_out) prec_out[nout] = rec.p;
      if (nfa_out) nfa_out[nout] = log_nfa;
    }
    nout++;
  }
  free(scaled); free(L.angles); free(L.modgrad); free(L.used); free(L.order); free(L.reg);
  return nout;
}

/* Stage dump for the per-stage GPU parity tests: the 0.8-scaled image, the
 * level-line angle in DEGREES (float, -1024 where undefined) and the pseudo-ordered
 * list restricted to defined pixels (the only ones that can seed a region). */
int orc_lsd_stages(const uint8_t* img, int w0, int h0, uint8_t* scaled_out, float* ang_deg_out,
                   int* order_out, int* ws_out, int* hs_out) {
  const double prec = ORC_PI * 22.5 / 180;
  const double rho = 2.0 / sin(prec);
  Lsd L;
  memset(&L, 0, sizeof(L));
  uint8_t* g = (uint8_t*)malloc((size_t)w0 * h0);
  orc_gaussian_blur7_s075(img, w0, h0, g);
  orc_resize_08(g, w0, h0, NULL, &L.w, &L.h);
  uint8_t* scaled = (uint8_t*)malloc((size_t)L.w * L.h);
  orc_resize_08(g, w0, h0, scaled, &L.w, &L.h);
  free(g);
  const size_t np = (size_t)L.w * L.h;
  L.angles = (double*)malloc(np * sizeof(double));
  L.modgrad = (double*)malloc(np * sizeof(double));
  L.order = (int*)malloc(np * sizeof(int));
  ll_angle(&L, scaled, rho, 1024);
  *ws_out = L.w;
  *hs_out = L.h;
  if (scaled_out) memcpy(scaled_out, scaled, np);
  if (ang_deg_out) {
    for (int y = 0; y < L.h; ++y)
      for (int x = 0; x < L.w; ++x) {
        float a = -1024.0f;
        if (L.angles[(size_t)y * L.w + x] != NOTDEF) {
          const uint8_t* r0 = scaled + (size_t)y * L.w;
          const uint8_t* r1 = r0 + L.w;
          int DA = r1[x + 1] - r0[x], BC = r0[x + 1] - r1[x];
          a = orc_fast_atan2((float)(DA + BC), (float)-(DA - BC));
        }
        ang_deg_out[(size_t)y * L.w + x] = a;
      }
  }
  int n = 0;
  for (int i = 0; i < L.n_order; ++i)
    if (L.angles[L.order[i]] != NOTDEF) {
      if (order_out) order_out[n] = L.order[i];
      n++;
    }
  free(scaled); free(L.angles); free(L.modgrad); free(L.order);
  return n;
}

/* Candidate dump (see g_cand_sink): runs LSD (ADV, 0.8 scaling) and returns the number of
 * candidates; out holds 16 doubles per candidate.  Not thread safe (test use only). */
int orc_lsd_candidates(const uint8_t* img, int w, int h, double* out, int cap) {
  float* seg = (float*)malloc((size_t)4 * 65536 * sizeof(float));
  g_cand_sink = out; g_cand_cap = cap; g_cand_n = 0;
  orc_lsd_detect(img, w, h, 2, 1, seg, NULL, NULL, NULL, 65536);
  g_cand_sink = NULL;
  free(seg);
  return g_cand_n;
}
