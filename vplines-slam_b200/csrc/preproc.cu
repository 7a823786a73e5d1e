// preproc.cu -- pre-processing in front of the line path (sm_100a): undistortion remap and
// CLAHE, so that raw frames are uploaded once and never leave HBM before detection.
//
// Replaces cv::remap(_img, img, undist_map1_, undist_map2_, CV_INTER_LINEAR) and
// cv::createCLAHE(3.0, Size(8,8))->apply(img, img) of LineFeatureTracker::readImage
// (/root/reference/feature_tracker/src/line_feature_tracker.cpp:62 and :64-68; the maps are the
// CV_32FC1 pair built by camera_model/src/camera_models/PinholeCamera.cc:729-790).  Integer /
// float32 arithmetic identical to OpenCV's (imgwarp.cpp remapBilinear + initInterTab2D, clahe.cpp);
// CPU restatement: oracle/orc_preproc.c, pinned against cv2 4.13.  All three kernels are
// HBM/L2-gather bound: remap reads 8 B of map + a 2x2 gather and writes 1 B per pixel; CLAHE
// reads the frame twice (histograms, interpolation) and writes it once.
#include "vpl_common.cuh"

namespace vpl {

// ---------------------------------------------------------------------------
// remap, INTER_LINEAR, BORDER_CONSTANT(0).  Maps are shared by all frames of the batch.
// sx = cvRound(mapx * 32): 5 fractional bits; the 2x2 weights come from OpenCV's 32x32 table
// (uint16, sum 32768); result = (sum + 2^14) >> 15.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
remap_kernel(const uint8_t* __restrict__ src_, uint8_t* __restrict__ dst_, const float* __restrict__ mapx,
             const float* __restrict__ mapy, const uint16_t* __restrict__ wtab, int w, int h) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const uint8_t* src = src_ + (size_t)blockIdx.z * w * h;
  const size_t o = (size_t)y * w + x;
  int sx = __float2int_rn(__fmul_rn(__ldg(mapx + o), 32.0f));
  int sy = __float2int_rn(__fmul_rn(__ldg(mapy + o), 32.0f));
  int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));
  const uint16_t* wv = wtab + (((sy & 31) << 5) + (sx & 31)) * 4;
  const ushort4 wq = *reinterpret_cast<const ushort4*>(wv);
  const bool x0 = (ix >= 0 && ix < w), x1 = (ix + 1 >= 0 && ix + 1 < w);
  const bool y0 = (iy >= 0 && iy < h), y1 = (iy + 1 >= 0 && iy + 1 < h);
  int acc = 0;
  if (y0) {
    const uint8_t* r = src + (size_t)iy * w;
    if (x0) acc += __ldg(r + ix) * (int)wq.x;
    if (x1) acc += __ldg(r + ix + 1) * (int)wq.y;
  }
  if (y1) {
    const uint8_t* r = src + (size_t)(iy + 1) * w;
    if (x0) acc += __ldg(r + ix) * (int)wq.z;
    if (x1) acc += __ldg(r + ix + 1) * (int)wq.w;
  }
  int v = (acc + (1 << 14)) >> 15;
  dst_[(size_t)blockIdx.z * w * h + o] = (uint8_t)max(0, min(255, v));
}

void launch_remap(const uint8_t* src, uint8_t* dst, const float* mapx, const float* mapy, const uint16_t* wtab,
                  int w, int h, int batch, cudaStream_t st) {
  dim3 grid((w + 31) / 32, (h + 7) / 8, batch);
  remap_kernel<<<grid, 256, 0, st>>>(src, dst, mapx, mapy, wtab, w, h);
}

// ---------------------------------------------------------------------------
// CLAHE step 1: one CTA per (tile, frame): histogram in shared memory, clip + redistribute,
// inclusive scan, LUT = saturate(cvRound(sum * 255/area)).  When the image size is not a
// multiple of the grid, OpenCV pads BOTH dimensions by tiles - (size % tiles) with
// BORDER_REFLECT_101; the tile pixels are read through the same reflection.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
clahe_lut_kernel(const uint8_t* __restrict__ src_, uint8_t* __restrict__ lut_, int w, int h, int tiles, int tw,
                 int th, int clip_limit, float lut_scale) {
  __shared__ int s_hist[256];
  __shared__ int s_scan[256];
  __shared__ int s_clipped;
  const int tile = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
  const int ty = tile / tiles, tx = tile - ty * tiles;
  const uint8_t* src = src_ + (size_t)f * w * h;
  s_hist[tid] = 0;
  if (tid == 0) s_clipped = 0;
  __syncthreads();
  const int area = tw * th;
  for (int i = tid; i < area; i += 256) {
    int yy = i / tw, xx = i - yy * tw;
    int gy = refl101(ty * th + yy, h), gx = refl101(tx * tw + xx, w);
    atomicAdd(&s_hist[__ldg(src + (size_t)gy * w + gx)], 1);
  }
  __syncthreads();
  int hv = s_hist[tid];
  if (clip_limit > 0) {
    if (hv > clip_limit) {
      atomicAdd(&s_clipped, hv - clip_limit);
      hv = clip_limit;
    }
    __syncthreads();
    const int clipped = s_clipped;
    const int redist = clipped / 256;
    int residual = clipped - redist * 256;
    hv += redist;
    if (residual != 0) {
      int step = max(256 / residual, 1);
      // for (i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++
      if (tid % step == 0 && tid / step < residual) hv += 1;
    }
  }
  // inclusive scan over the 256 bins
  s_scan[tid] = hv;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    int t = (tid >= o) ? s_scan[tid - o] : 0;
    __syncthreads();
    s_scan[tid] += t;
    __syncthreads();
  }
  int r = __float2int_rn(__fmul_rn((float)s_scan[tid], lut_scale));
  lut_[((size_t)f * tiles * tiles + tile) * 256 + tid] = (uint8_t)max(0, min(255, r));
}

// CLAHE step 2: bilinear interpolation between the four surrounding tile LUTs (float32, unfused).
__global__ void __launch_bounds__(256)
clahe_apply_kernel(const uint8_t* __restrict__ src_, const uint8_t* __restrict__ lut_, uint8_t* __restrict__ dst_,
                   int w, int h, int tiles, float inv_tw, float inv_th) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const size_t fo = (size_t)blockIdx.z * w * h;
  const uint8_t* lut = lut_ + (size_t)blockIdx.z * tiles * tiles * 256;
  float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
  int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
  float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
  ty1 = max(ty1, 0);
  ty2 = min(ty2, tiles - 1);
  float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
  int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
  float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
  tx1 = max(tx1, 0);
  tx2 = min(tx2, tiles - 1);
  const int v = __ldg(src_ + fo + (size_t)y * w + x);
  const uint8_t* p1 = lut + (size_t)ty1 * tiles * 256;
  const uint8_t* p2 = lut + (size_t)ty2 * tiles * 256;
  float a = __fadd_rn(__fmul_rn((float)__ldg(p1 + tx1 * 256 + v), xa1), __fmul_rn((float)__ldg(p1 + tx2 * 256 + v), xa));
  float b = __fadd_rn(__fmul_rn((float)__ldg(p2 + tx1 * 256 + v), xa1), __fmul_rn((float)__ldg(p2 + tx2 * 256 + v), xa));
  float res = __fadd_rn(__fmul_rn(a, ya1), __fmul_rn(b, ya));
  int r = __float2int_rn(res);
  dst_[fo + (size_t)y * w + x] = (uint8_t)max(0, min(255, r));
}

void launch_clahe(const uint8_t* src, uint8_t* dst, uint8_t* lut, int w, int h, double clip, int tiles, int batch,
                  cudaStream_t st) {
  int wp = w, hp = h;
  if (!(w % tiles == 0 && h % tiles == 0)) {
    wp = w + (tiles - (w % tiles));
    hp = h + (tiles - (h % tiles));
  }
  const int tw = wp / tiles, th = hp / tiles;
  const int area = tw * th;
  int clip_limit = 0;
  if (clip > 0.0) {
    clip_limit = (int)(clip * area / 256);
    if (clip_limit < 1) clip_limit = 1;
  }
  const float lut_scale = (float)255 / (float)area;
  clahe_lut_kernel<<<dim3(tiles * tiles, batch), 256, 0, st>>>(src, lut, w, h, tiles, tw, th, clip_limit, lut_scale);
  dim3 grid((w + 31) / 32, (h + 7) / 8, batch);
  clahe_apply_kernel<<<grid, 256, 0, st>>>(src, lut, dst, w, h, tiles, 1.0f / (float)tw, 1.0f / (float)th);
}

}  // namespace vpl
