/*
 * orc_lbd.c -- CPU oracle: LSDDetector::detect KeyLine packing, BinaryDescriptor::
 * compute (LBD) and brute-force Hamming kNN.  TEST INFRASTRUCTURE ONLY (see
 * vpl_oracle.h).
 *
 * Restates opencv_contrib 3.4 modules/line_descriptor/src/{LSDDetector.cpp,
 * binary_descriptor.cpp, binary_descriptor_matcher.cpp} (third party, NOT in
 * /root/reference and not installed: cv2 here has no line_descriptor module).
 * PARITY UNPINNED for KeyLine packing and LBD: no upstream vector exists in the
 * reference; this file is the definition the CUDA
 * path is held to, bit for bit.  Hamming kNN follows SURVEY Appendix C (brute
 * force, lowest train index on ties) and is pinned against cv2.BFMatcher.
 *
 * Choices where the published code is ambiguous (all documented in DESIGN.md):
 *  - LSDDetector blurs octave 0 with GaussianBlur(5x5, sigma 1) when blur_first
 *    (SURVEY N1); blur_first=0 gives the variant with that line commented out.
 *  - gaussCoefL_: the published code computes sigma=(widthOfBand_*2+1)/2 and
 *    u=(widthOfBand_*3-1)/2 in INTEGER arithmetic => sigma=7, u=10; likewise the
 *    global table u=sigma=31.
 *  - float math is plain IEEE float32, no fused multiply-add, samples summed
 *    left to right along a row, rows top to bottom.
 *  - cos/sin/atan2/sqrt of a float resolve to the C++ float overloads: the
 *    correctly rounded float result (sqrtf; (float)cos((double)x) etc.).
 */
#include "vpl_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NUM_OF_BANDS 9
#define WIDTH_OF_BAND 7

/* ---- LSDDetector::detect ------------------------------------------------- */
static void check_line_extremes(float* e, int cols, int rows) {
  if (e[0] < 0) e[0] = 0;
  if (e[0] >= cols) e[0] = (float)cols - 1.0f;
  if (e[2] < 0) e[2] = 0;
  if (e[2] >= cols) e[2] = (float)cols - 1.0f;
  if (e[1] < 0) e[1] = 0;
  if (e[1] >= rows) e[1] = (float)rows - 1.0f;
  if (e[3] < 0) e[3] = 0;
  if (e[3] >= rows) e[3] = (float)rows - 1.0f;
}

/* cv::LineIterator(img, Point(p1), Point(p2), 8).count with both points inside
 * the image: max(|dx|,|dy|)+1 on cvRound()ed (half-to-even) coordinates. */
static int line_iterator_count(const float* e) {
  int x0 = (int)lrintf(e[0]), y0 = (int)lrintf(e[1]);
  int x1 = (int)lrintf(e[2]), y1 = (int)lrintf(e[3]);
  int dx = abs(x1 - x0), dy = abs(y1 - y0);
  return (dx > dy ? dx : dy) + 1;
}

int orc_lsd_detector_detect(const uint8_t* img, int w, int h, int scale, int num_octaves,
                            int blur_first, OrcKeyLine* out, int cap) {
  uint8_t* cur = (uint8_t*)malloc((size_t)w * h);
  if (blur_first) orc_gaussian_blur5(img, w, h, cur);
  else memcpy(cur, img, (size_t)w * h);
  int cw = w, ch = h;
  int class_counter = -1;
  int nout = 0;
  const int seg_cap = 1 << 16;
  float* seg = (float*)malloc((size_t)seg_cap * 4 * sizeof(float));
  for (int oct = 0; oct < num_octaves; ++oct) {
    if (oct > 0) {
      /* pyrDown(cur, cur, Size(cols/scale, rows/scale)); only scale==2 is a
       * valid pyrDown destination size. */
      uint8_t* nxt = (uint8_t*)malloc((size_t)(cw / 2) * (ch / 2));
      orc_pyrdown_half(cur, cw, ch, nxt);
      free(cur);
      cur = nxt;
      cw /= 2;
      ch /= 2;
    }
    int n = orc_lsd_detect(cur, cw, ch, 2 /*ADV*/, 1, seg, NULL, NULL, NULL, seg_cap);
    if (n > seg_cap) n = seg_cap;
    float octaveScale = (float)pow((double)(float)scale, (double)oct);
    for (int k = 0; k < n; ++k) {
      float e[4] = {seg[4 * k], seg[4 * k + 1], seg[4 * k + 2], seg[4 * k + 3]};
      check_line_extremes(e, cw, ch);
      OrcKeyLine kl;
      kl.startPointX = e[0] * octaveScale;
      kl.startPointY = e[1] * octaveScale;
      kl.endPointX = e[2] * octaveScale;
      kl.endPointY = e[3] * octaveScale;
      kl.sPointInOctaveX = e[0];
      kl.sPointInOctaveY = e[1];
      kl.ePointInOctaveX = e[2];
      kl.ePointInOctaveY = e[3];
      {
        double a = (double)(float)(e[0] - e[2]), b = (double)(float)(e[1] - e[3]);
        kl.lineLength = (float)sqrt(a * a + b * b);
      }
      kl.numOfPixels = line_iterator_count(e);
      kl.angle = (float)atan2((double)(float)(kl.endPointY - kl.startPointY),
                              (double)(float)(kl.endPointX - kl.startPointX));
      kl.class_id = ++class_counter;
      kl.octave = oct;
      kl.size = (kl.endPointX - kl.startPointX) * (kl.endPointY - kl.startPointY);
      kl.response = kl.lineLength / (float)(cw > ch ? cw : ch);
      kl.pt_x = (kl.endPointX + kl.startPointX) / 2;
      kl.pt_y = (kl.endPointY + kl.startPointY) / 2;
      if (nout < cap) out[nout] = kl;
      nout++;
    }
  }
  free(seg);
  free(cur);
  return nout;
}

/* ---- BinaryDescriptor::compute ------------------------------------------- */
static const int combinations[32][2] = {
    {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
    {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
    {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

static void gauss_tables(double* coefL, double* coefG) {
  /* integer arithmetic exactly as published (see header comment) */
  double u = (WIDTH_OF_BAND * 3 - 1) / 2;
  double sigma = (WIDTH_OF_BAND * 2 + 1) / 2;
  double invsigma2 = -1 / (2 * sigma * sigma);
  for (int i = 0; i < WIDTH_OF_BAND * 3; ++i) {
    double dis = i - u;
    coefL[i] = exp(dis * dis * invsigma2);
  }
  u = (NUM_OF_BANDS * WIDTH_OF_BAND - 1) / 2;
  sigma = u;
  invsigma2 = -1 / (2 * sigma * sigma);
  for (int i = 0; i < NUM_OF_BANDS * WIDTH_OF_BAND; ++i) {
    double dis = i - u;
    coefG[i] = exp(dis * dis * invsigma2);
  }
}

static void lbd_one(const int16_t* pdx, const int16_t* pdy, int realWidth, int realHeight,
                    const OrcKeyLine* kl, const double* gaussCoefL, const double* gaussCoefG,
                    float* desVec /*72*/) {
  const short heightOfLSP = WIDTH_OF_BAND * NUM_OF_BANDS;
  float bandSum[8][NUM_OF_BANDS]; /* pgdL ngdL pgdL2 ngdL2 pgdO ngdO pgdO2 ngdO2 */
  memset(bandSum, 0, sizeof(bandSum));
  const short imageWidth = (short)(realWidth - 1), imageHeight = (short)(realHeight - 1);
  const short lengthOfLSP = (short)kl->numOfPixels;
  const short halfHeight = (heightOfLSP - 1) / 2;
  const short halfWidth = (lengthOfLSP - 1) / 2;
  const float lineMiddlePointX = (float)(0.5 * (kl->sPointInOctaveX + kl->ePointInOctaveX));
  const float lineMiddlePointY = (float)(0.5 * (kl->sPointInOctaveY + kl->ePointInOctaveY));
  float dL[2], dO[2];
  dL[0] = (float)cos((double)kl->angle);
  dL[1] = (float)sin((double)kl->angle);
  dO[0] = -dL[1];
  dO[1] = dL[0];
  float sCorX0 = -dL[0] * halfWidth + dL[1] * halfHeight + lineMiddlePointX;
  float sCorY0 = -dL[1] * halfWidth - dL[0] * halfHeight + lineMiddlePointY;
  for (short hID = 0; hID < heightOfLSP; hID++) {
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdLRowSum = 0, ngdLRowSum = 0, pgdORowSum = 0, ngdORowSum = 0;
    for (short wID = 0; wID < lengthOfLSP; wID++) {
      short tempCor = (short)roundf(sCorX);
      short xCor = (tempCor < 0) ? 0 : (tempCor > imageWidth) ? imageWidth : tempCor;
      tempCor = (short)roundf(sCorY);
      short yCor = (tempCor < 0) ? 0 : (tempCor > imageHeight) ? imageHeight : tempCor;
      short dx = pdx[yCor * realWidth + xCor];
      short dy = pdy[yCor * realWidth + xCor];
      float gDL = dx * dL[0] + dy * dL[1];
      float gDO = dx * dO[0] + dy * dO[1];
      if (gDL > 0) pgdLRowSum += gDL;
      else ngdLRowSum -= gDL;
      if (gDO > 0) pgdORowSum += gDO;
      else ngdORowSum -= gDO;
      sCorX += dL[0];
      sCorY += dL[1];
    }
    sCorX0 -= dL[1];
    sCorY0 += dL[0];
    float coefInGaussion = (float)gaussCoefG[hID];
    pgdLRowSum = coefInGaussion * pgdLRowSum;
    ngdLRowSum = coefInGaussion * ngdLRowSum;
    float pgdL2RowSum = pgdLRowSum * pgdLRowSum;
    float ngdL2RowSum = ngdLRowSum * ngdLRowSum;
    pgdORowSum = coefInGaussion * pgdORowSum;
    ngdORowSum = coefInGaussion * ngdORowSum;
    float pgdO2RowSum = pgdORowSum * pgdORowSum;
    float ngdO2RowSum = ngdORowSum * ngdORowSum;
    const float rs[8] = {pgdLRowSum, ngdLRowSum, pgdL2RowSum, ngdL2RowSum,
                         pgdORowSum, ngdORowSum, pgdO2RowSum, ngdO2RowSum};
    short bandID = (short)(hID / WIDTH_OF_BAND);
    for (int pass = 0; pass < 3; ++pass) {
      int b, ci;
      if (pass == 0) { b = bandID; ci = hID % WIDTH_OF_BAND + WIDTH_OF_BAND; }
      else if (pass == 1) { b = bandID - 1; ci = hID % WIDTH_OF_BAND + 2 * WIDTH_OF_BAND; }
      else { b = bandID + 1; ci = hID % WIDTH_OF_BAND; }
      if (b < 0 || b >= NUM_OF_BANDS) continue;
      coefInGaussion = (float)gaussCoefL[ci];
      for (int q = 0; q < 8; ++q) {
        if (q == 2 || q == 3 || q == 6 || q == 7) bandSum[q][b] += coefInGaussion * coefInGaussion * rs[q];
        else bandSum[q][b] += coefInGaussion * rs[q];
      }
    }
  }
  const float invN2 = (float)(1.0 / (WIDTH_OF_BAND * 2.0));
  const float invN3 = (float)(1.0 / (WIDTH_OF_BAND * 3.0));
  for (int bandID = 0; bandID < NUM_OF_BANDS; bandID++) {
    float invN = (bandID == 0 || bandID == NUM_OF_BANDS - 1) ? invN2 : invN3;
    int desID = bandID * 8;
    float temp = bandSum[0][bandID] * invN;
    desVec[desID] = temp;
    desVec[desID + 4] = sqrtf(bandSum[2][bandID] * invN - temp * temp);
    temp = bandSum[1][bandID] * invN;
    desVec[desID + 1] = temp;
    desVec[desID + 5] = sqrtf(bandSum[3][bandID] * invN - temp * temp);
    temp = bandSum[4][bandID] * invN;
    desVec[desID + 2] = temp;
    desVec[desID + 6] = sqrtf(bandSum[6][bandID] * invN - temp * temp);
    temp = bandSum[5][bandID] * invN;
    desVec[desID + 3] = temp;
    desVec[desID + 7] = sqrtf(bandSum[7][bandID] * invN - temp * temp);
  }
  float tempM = 0, tempS = 0;
  for (int b = 0; b < NUM_OF_BANDS; ++b) {
    const float* d = desVec + 8 * b;
    tempM += d[0] * d[0]; tempM += d[1] * d[1]; tempM += d[2] * d[2]; tempM += d[3] * d[3];
    tempS += d[4] * d[4]; tempS += d[5] * d[5]; tempS += d[6] * d[6]; tempS += d[7] * d[7];
  }
  tempM = 1.0f / sqrtf(tempM);
  tempS = 1.0f / sqrtf(tempS);
  for (int b = 0; b < NUM_OF_BANDS; ++b) {
    float* d = desVec + 8 * b;
    d[0] = d[0] * tempM; d[1] = d[1] * tempM; d[2] = d[2] * tempM; d[3] = d[3] * tempM;
    d[4] = d[4] * tempS; d[5] = d[5] * tempS; d[6] = d[6] * tempS; d[7] = d[7] * tempS;
  }
  for (int i = 0; i < 72; ++i)
    if ((double)desVec[i] > 0.4) desVec[i] = (float)0.4;
  float temp = 0;
  for (int i = 0; i < 72; ++i) temp += desVec[i] * desVec[i];
  temp = 1.0f / sqrtf(temp);
  for (int i = 0; i < 72; ++i) desVec[i] = desVec[i] * temp;
}

int orc_lbd_compute(const uint8_t* img, int w, int h, const OrcKeyLine* kl, int n, uint8_t* desc,
                    float* fdesc) {
  if (n <= 0) return 0;
  int maxOct = 0;
  for (int i = 0; i < n; ++i)
    if (kl[i].octave > maxOct) maxOct = kl[i].octave;
  int nOct = maxOct + 1;
  double coefL[WIDTH_OF_BAND * 3], coefG[NUM_OF_BANDS * WIDTH_OF_BAND];
  gauss_tables(coefL, coefG);
  int16_t** dxs = (int16_t**)calloc((size_t)nOct, sizeof(int16_t*));
  int16_t** dys = (int16_t**)calloc((size_t)nOct, sizeof(int16_t*));
  int* ws = (int*)calloc((size_t)nOct, sizeof(int));
  int* hs = (int*)calloc((size_t)nOct, sizeof(int));
  uint8_t* cur = (uint8_t*)malloc((size_t)w * h);
  orc_gaussian_blur5(img, w, h, cur);
  int cw = w, ch = h;
  for (int o = 0; o < nOct; ++o) {
    if (o > 0) {
      uint8_t* nxt = (uint8_t*)malloc((size_t)(cw / 2) * (ch / 2));
      orc_pyrdown_half(cur, cw, ch, nxt);
      free(cur);
      cur = nxt;
      cw /= 2;
      ch /= 2;
    }
    ws[o] = cw;
    hs[o] = ch;
    dxs[o] = (int16_t*)malloc((size_t)cw * ch * sizeof(int16_t));
    dys[o] = (int16_t*)malloc((size_t)cw * ch * sizeof(int16_t));
    orc_sobel3(cur, cw, ch, dxs[o], dys[o]);
  }
  free(cur);
  for (int i = 0; i < n; ++i) {
    float d[72];
    int o = kl[i].octave;
    lbd_one(dxs[o], dys[o], ws[o], hs[o], &kl[i], coefL, coefG, d);
    if (fdesc) memcpy(fdesc + (size_t)72 * i, d, sizeof(d));
    for (int c = 0; c < 32; ++c) {
      const float* f1 = d + 8 * combinations[c][0];
      const float* f2 = d + 8 * combinations[c][1];
      uint8_t r = 0;
      for (int b = 0; b < 8; ++b)
        if (f1[b] > f2[b]) r += (uint8_t)(1 << b);
      desc[(size_t)32 * i + c] = r;
    }
  }
  for (int o = 0; o < nOct; ++o) { free(dxs[o]); free(dys[o]); }
  free(dxs); free(dys); free(ws); free(hs);
  return n;
}

/* ---- brute-force Hamming kNN --------------------------------------------- */
void orc_hamming_knn(const uint8_t* q, int nq, const uint8_t* t, int nt, int k, int32_t* idx,
                     int32_t* dist) {
  int* d = (int*)malloc((size_t)(nt > 0 ? nt : 1) * sizeof(int));
  for (int i = 0; i < nq; ++i) {
    for (int j = 0; j < nt; ++j) {
      int s = 0;
      for (int b = 0; b < 32; ++b) s += __builtin_popcount((unsigned)(q[32 * i + b] ^ t[32 * j + b]));
      d[j] = s;
    }
    /* k passes of "smallest (dist, index) strictly after the previous pick" */
    int pd = -1, pj = -1;
    for (int r = 0; r < k; ++r) {
      int bd = 1 << 30, bj = -1;
      for (int j = 0; j < nt; ++j) {
        if (d[j] < pd || (d[j] == pd && j <= pj)) continue;
        if (d[j] < bd) { bd = d[j]; bj = j; }
      }
      idx[(size_t)i * k + r] = bj;
      dist[(size_t)i * k + r] = bj >= 0 ? bd : -1;
      if (bj < 0) { pd = 1 << 30; pj = 1 << 30; }
      else { pd = bd; pj = bj; }
    }
  }
  free(d);
}

/* ---- whole front end on a sequence (CPU baseline timing) ------------------ */
int64_t orc_frontend_sequence(const uint8_t* frames, int n_frames, int w, int h, int num_octaves,
                              int max_lines) {
  OrcKeyLine* kl = (OrcKeyLine*)malloc((size_t)max_lines * sizeof(OrcKeyLine));
  uint8_t* dcur = (uint8_t*)malloc((size_t)max_lines * 32);
  uint8_t* dprev = (uint8_t*)malloc((size_t)max_lines * 32);
  int32_t* mi = (int32_t*)malloc((size_t)max_lines * sizeof(int32_t));
  int32_t* md = (int32_t*)malloc((size_t)max_lines * sizeof(int32_t));
  int nprev = 0;
  int64_t total = 0, chk = 0;
  for (int f = 0; f < n_frames; ++f) {
    const uint8_t* img = frames + (size_t)f * w * h;
    int n = orc_lsd_detector_detect(img, w, h, 2, num_octaves, 1, kl, max_lines);
    if (n > max_lines) n = max_lines;
    orc_lbd_compute(img, w, h, kl, n, dcur, NULL);
    if (f > 0 && n > 0 && nprev > 0) {
      orc_hamming_knn(dcur, n, dprev, nprev, 1, mi, md);
      for (int i = 0; i < n; ++i) chk += mi[i] + md[i];
    }
    uint8_t* t = dprev; dprev = dcur; dcur = t;
    nprev = n;
    total += n;
  }
  free(kl); free(dcur); free(dprev); free(mi); free(md);
  return total + (chk & 0);
}
