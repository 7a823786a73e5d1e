/*
 * orc_preproc.c -- CPU oracle of the pre-processing in front of the line path: cv::remap
 * (undistortion, INTER_LINEAR, BORDER_CONSTANT 0, CV_32FC1 maps) and cv::CLAHE.  TEST
 * INFRASTRUCTURE ONLY (see vpl_oracle.h).
 *
 * What it follows: LineFeatureTracker::readImage, /root/reference/feature_tracker/src/
 * line_feature_tracker.cpp:62 (cv::remap with undist_map1_/undist_map2_, the CV_32FC1 maps of
 * camera_model/src/camera_models/PinholeCamera.cc:729-790) and :64-68 (createCLAHE(3.0, Size(8,8))
 * when EQUALIZE).  The arithmetic is OpenCV's (imgproc imgwarp.cpp remapBilinear + initInterTab2D,
 * clahe.cpp), un-vendored; pinned bit-for-bit against cv2 4.13 by tests/test_oracle_preproc.py.
 */
#include "vpl_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define INTER_BITS 5
#define INTER_TAB_SIZE 32
#define INTER_REMAP_COEF_SCALE 32768

/* initInterTab2D(INTER_LINEAR, fixpt): 32x32 entries of 4 weights summing to 32768.  The entry
 * for (fx,fy) = (0,0) is 32768 itself (cv2 4.13 behaves as if the table were unsigned). */
void orc_remap_weight_table(uint16_t* tab /* 1024 x 4 */) {
  for (int i = 0; i < INTER_TAB_SIZE; ++i) {
    float fy = (float)i / (float)INTER_TAB_SIZE;
    for (int j = 0; j < INTER_TAB_SIZE; ++j) {
      float fx = (float)j / (float)INTER_TAB_SIZE;
      float wy[2] = {1.0f - fy, fy}, wx[2] = {1.0f - fx, fx};
      int it[4], isum = 0;
      for (int k1 = 0; k1 < 2; ++k1)
        for (int k2 = 0; k2 < 2; ++k2) {
          float v = wy[k1] * wx[k2];
          long r = lrintf(v * (float)INTER_REMAP_COEF_SCALE);
          if (r > 32767) r = 32767;
          if (r < -32768) r = -32768;
          it[k1 * 2 + k2] = (int)r;
          isum += (int)r;
        }
      if (isum != INTER_REMAP_COEF_SCALE) {
        int diff = isum - INTER_REMAP_COEF_SCALE;
        int kmax = 0, kmin = 0;
        for (int k = 1; k < 4; ++k) {
          if (it[k] > it[kmax]) kmax = k;
          if (it[k] < it[kmin]) kmin = k;
        }
        if (diff < 0) it[kmax] -= diff;
        else it[kmin] -= diff;
      }
      for (int k = 0; k < 4; ++k) tab[(i * INTER_TAB_SIZE + j) * 4 + k] = (uint16_t)it[k];
    }
  }
}

void orc_remap_linear(const uint8_t* src, int w, int h, const float* mapx, const float* mapy, int dw, int dh,
                      uint8_t* dst) {
  static uint16_t tab[INTER_TAB_SIZE * INTER_TAB_SIZE * 4];
  static int have = 0;
  if (!have) { orc_remap_weight_table(tab); have = 1; }
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      int sx = (int)lrintf(mapx[(size_t)y * dw + x] * (float)INTER_TAB_SIZE);
      int sy = (int)lrintf(mapy[(size_t)y * dw + x] * (float)INTER_TAB_SIZE);
      int ix = sx >> INTER_BITS, iy = sy >> INTER_BITS;
      /* the block converter stores the integer part as short (saturate_cast<short>) */
      if (ix > 32767) ix = 32767;
      if (ix < -32768) ix = -32768;
      if (iy > 32767) iy = 32767;
      if (iy < -32768) iy = -32768;
      const uint16_t* wv = tab + (((sy & (INTER_TAB_SIZE - 1)) * INTER_TAB_SIZE) + (sx & (INTER_TAB_SIZE - 1))) * 4;
      int acc = 0;
      for (int k = 0; k < 4; ++k) {
        int yy = iy + (k >> 1), xx = ix + (k & 1);
        int p = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? src[(size_t)yy * w + xx] : 0;
        acc += p * (int)wv[k];
      }
      int v = (acc + (1 << 14)) >> 15;
      dst[(size_t)y * dw + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
}

static inline int refl101p(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * n - 2 - i;
  }
  return i;
}

/* cv::CLAHE::apply on CV_8UC1 (clahe.cpp): tiles x tiles grid. */
void orc_clahe(const uint8_t* src, int w, int h, double clip_limit, int tiles, uint8_t* dst) {
  int wp = w, hp = h;
  if (!(w % tiles == 0 && h % tiles == 0)) {  /* cv2 pads BOTH sizes by tiles - (size % tiles) */
    wp = w + (tiles - (w % tiles));
    hp = h + (tiles - (h % tiles));
  }
  const int tw = wp / tiles, th = hp / tiles;
  const int area = tw * th;
  const float lutScale = (float)255 / (float)area;
  int clipLimit = 0;
  if (clip_limit > 0.0) {
    clipLimit = (int)(clip_limit * area / 256);
    if (clipLimit < 1) clipLimit = 1;
  }
  uint8_t* lut = (uint8_t*)malloc((size_t)tiles * tiles * 256);
  for (int ty = 0; ty < tiles; ++ty)
    for (int tx = 0; tx < tiles; ++tx) {
      int hist[256];
      memset(hist, 0, sizeof(hist));
      for (int y = ty * th; y < (ty + 1) * th; ++y)
        for (int x = tx * tw; x < (tx + 1) * tw; ++x)
          hist[src[(size_t)refl101p(y, h) * w + refl101p(x, w)]]++;
      if (clipLimit > 0) {
        int clipped = 0;
        for (int i = 0; i < 256; ++i)
          if (hist[i] > clipLimit) { clipped += hist[i] - clipLimit; hist[i] = clipLimit; }
        int redistBatch = clipped / 256;
        int residual = clipped - redistBatch * 256;
        for (int i = 0; i < 256; ++i) hist[i] += redistBatch;
        if (residual != 0) {
          int residualStep = 256 / residual;
          if (residualStep < 1) residualStep = 1;
          for (int i = 0; i < 256 && residual > 0; i += residualStep, residual--) hist[i]++;
        }
      }
      int sum = 0;
      uint8_t* L = lut + ((size_t)ty * tiles + tx) * 256;
      for (int i = 0; i < 256; ++i) {
        sum += hist[i];
        long r = lrintf((float)sum * lutScale);
        L[i] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
      }
    }
  const float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
  for (int y = 0; y < h; ++y) {
    float tyf = (float)y * inv_th - 0.5f;
    int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
    float ya = tyf - (float)ty1, ya1 = 1.0f - ya;
    if (ty1 < 0) ty1 = 0;
    if (ty2 > tiles - 1) ty2 = tiles - 1;
    for (int x = 0; x < w; ++x) {
      float txf = (float)x * inv_tw - 0.5f;
      int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
      float xa = txf - (float)tx1, xa1 = 1.0f - xa;
      if (tx1 < 0) tx1 = 0;
      if (tx2 > tiles - 1) tx2 = tiles - 1;
      int v = src[(size_t)y * w + x];
      const uint8_t* p1 = lut + ((size_t)ty1 * tiles) * 256;
      const uint8_t* p2 = lut + ((size_t)ty2 * tiles) * 256;
      float res = ((float)p1[tx1 * 256 + v] * xa1 + (float)p1[tx2 * 256 + v] * xa) * ya1 +
                  ((float)p2[tx1 * 256 + v] * xa1 + (float)p2[tx2 * 256 + v] * xa) * ya;
      long r = lrintf(res);
      dst[(size_t)y * w + x] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
    }
  }
  free(lut);
}
