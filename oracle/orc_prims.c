/*
 * orc_prims.c -- CPU oracle, image primitives.  TEST INFRASTRUCTURE ONLY (see
 * vpl_oracle.h).  Integer restatements of the OpenCV imgproc routines that
 * LSDDetector / BinaryDescriptor / LSD call (opencv 3.4 imgproc: smooth.cpp,
 * pyramids.cpp, resize.cpp, deriv.cpp -- not vendored in /root/reference;
 * recipes from SURVEY.md Appendix F, each pinned bit-exact against cv2 4.13 by
 * tests/test_oracle_prims.py).
 */
#include "vpl_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba */
static inline int refl101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * n - 2 - i;
  }
  return i;
}

/* Separable 8-bit fixed-point Gaussian, kernel taps sum to 256; horizontal pass
 * exact in 8.8, vertical in 16.16, one rounding: (v + 32768) >> 16. */
static void blur_fixed(const uint8_t* src, int w, int h, uint8_t* dst, const int* k, int r) {
  uint32_t* tmp = (uint32_t*)malloc((size_t)w * h * sizeof(uint32_t));
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      uint32_t s = 0;
      for (int i = -r; i <= r; ++i) s += (uint32_t)k[i + r] * src[(size_t)y * w + refl101(x + i, w)];
      tmp[(size_t)y * w + x] = s;
    }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      uint32_t s = 0;
      for (int i = -r; i <= r; ++i) s += (uint32_t)k[i + r] * tmp[(size_t)refl101(y + i, h) * w + x];
      dst[(size_t)y * w + x] = (uint8_t)((s + 32768u) >> 16);
    }
  free(tmp);
}

void orc_gaussian_blur5(const uint8_t* src, int w, int h, uint8_t* dst) {
  static const int k[5] = {14, 62, 104, 62, 14};
  blur_fixed(src, w, h, dst, k, 2);
}

void orc_gaussian_blur7_s075(const uint8_t* src, int w, int h, uint8_t* dst) {
  static const int k[7] = {0, 4, 56, 136, 56, 4, 0};
  blur_fixed(src, w, h, dst, k, 3);
}

/* resize(fx=fy=0.8, INTER_LINEAR_EXACT): source coord (d+0.5)*1.25-0.5, weights
 * are multiples of 1/8, a single round-half-up at the end. */
void orc_resize_08(const uint8_t* src, int w, int h, uint8_t* dst, int* dw_, int* dh_) {
  int dw = (int)lrint(w * 0.8), dh = (int)lrint(h * 0.8);
  *dw_ = dw;
  *dh_ = dh;
  if (!dst) return;
  for (int y = 0; y < dh; ++y) {
    /* 8*f = 10*d + 1  (f = (d+0.5)*1.25-0.5 = 1.25 d + 0.125) */
    int fy8 = 10 * y + 1;
    int iy = fy8 >> 3, ay = fy8 & 7;
    if (iy >= h - 1) { iy = h - 1; ay = 0; }
    int iy1 = iy + 1 < h ? iy + 1 : iy;
    for (int x = 0; x < dw; ++x) {
      int fx8 = 10 * x + 1;
      int ix = fx8 >> 3, ax = fx8 & 7;
      if (ix >= w - 1) { ix = w - 1; ax = 0; }
      int ix1 = ix + 1 < w ? ix + 1 : ix;
      int s = (8 - ay) * ((8 - ax) * src[(size_t)iy * w + ix] + ax * src[(size_t)iy * w + ix1]) +
              ay * ((8 - ax) * src[(size_t)iy1 * w + ix] + ax * src[(size_t)iy1 * w + ix1]);
      dst[(size_t)y * dw + x] = (uint8_t)((s + 32) >> 6);
    }
  }
}

/* pyrDown(src, dst, Size(w/2, h/2)): [1 4 6 4 1]x[1 4 6 4 1] centred on (2x,2y),
 * REFLECT_101, (sum+128)>>8. */
void orc_pyrdown_half(const uint8_t* src, int w, int h, uint8_t* dst) {
  static const int k[5] = {1, 4, 6, 4, 1};
  int dw = w / 2, dh = h / 2;
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      int s = 0;
      for (int j = -2; j <= 2; ++j) {
        int yy = refl101(2 * y + j, h);
        int rs = 0;
        for (int i = -2; i <= 2; ++i) rs += k[i + 2] * src[(size_t)yy * w + refl101(2 * x + i, w)];
        s += k[j + 2] * rs;
      }
      dst[(size_t)y * dw + x] = (uint8_t)((s + 128) >> 8);
    }
}

/* Sobel 3x3, CV_16S, no scaling, REFLECT_101. */
void orc_sobel3(const uint8_t* src, int w, int h, int16_t* dx, int16_t* dy) {
  for (int y = 0; y < h; ++y) {
    int ym = refl101(y - 1, h), yp = refl101(y + 1, h);
    for (int x = 0; x < w; ++x) {
      int xm = refl101(x - 1, w), xp = refl101(x + 1, w);
#define P(yy, xx) ((int)src[(size_t)(yy) * w + (xx)])
      int gx = (P(ym, xp) + 2 * P(y, xp) + P(yp, xp)) - (P(ym, xm) + 2 * P(y, xm) + P(yp, xm));
      int gy = (P(yp, xm) + 2 * P(yp, x) + P(yp, xp)) - (P(ym, xm) + 2 * P(ym, x) + P(ym, xp));
#undef P
      dx[(size_t)y * w + x] = (int16_t)gx;
      dy[(size_t)y * w + x] = (int16_t)gy;
    }
  }
}

/* cv::fastAtan2 scalar path (opencv core mathfuncs_core: atan_f32), degrees in
 * [0,360); all float32, no fused multiply-add (build with -ffp-contract=off). */
float orc_fast_atan2(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale;
  const float p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale;
  const float p7 = -0.04432655554792128f * scale;
  float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}
