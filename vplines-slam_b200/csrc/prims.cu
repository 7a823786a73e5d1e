// prims.cu -- streaming image kernels of the line front end (sm_100a).
//
// Integer arithmetic only, bit-identical to OpenCV's GaussianBlur / pyrDown /
// Sobel / resize(INTER_LINEAR_EXACT) (the CPU restatement
// they are tested against is oracle/orc_prims.c).  They replace the image passes
// that LSDDetector::detect and BinaryDescriptor::compute run per frame
// (opencv_contrib 3.4 LSDDetector.cpp computeGaussianPyramid, binary_descriptor.cpp
// computeGaussianPyramid/computeSobel) and LSD's pre-scaling (lsd.cpp flsd).
// All are HBM-bound: shared-memory halo tiles, 32-bit/128-bit global accesses,
// one grid covering the whole batch.
#include "vpl_common.cuh"

namespace vpl {

// ---------------------------------------------------------------------------
// K1: GaussianBlur(5x5, sigma 1) fused with Sobel 3x3 (dx,dy int16).
// Fixed-point kernel [14 62 104 62 14]/256: horizontal pass exact in 8.8,
// vertical in 16.16, (v + 32768) >> 16.  BORDER_REFLECT_101 on both stages; the
// reflect-extended input is symmetric about the border pixel and the kernel is
// symmetric, so blurring the extended tile yields exactly the reflected blurred
// halo the Sobel stage needs.
// Algorithmic bytes: read P, write P (pyr) + 4P (grad) = 6P per frame.
// ---------------------------------------------------------------------------
constexpr int B5_TW = 64, B5_TH = 32, B5_THREADS = 256;
constexpr int B5_IW = B5_TW + 8;  // input tile: x0-4 .. x0+TW+4 (word aligned)
constexpr int B5_IH = B5_TH + 6;  // y0-3 .. y0+TH+3

__global__ void __launch_bounds__(B5_THREADS)
blur5_sobel_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ pyr,
                   short2* __restrict__ grad, int w, int h, int do_blur) {
  __shared__ __align__(16) uint8_t s_in[B5_IH][B5_IW];
  __shared__ uint16_t s_hb[B5_IH][B5_TW + 2];
  __shared__ __align__(4) uint8_t s_bl[B5_TH + 2][B5_TW + 4];  // [.][1 + x], x=-1..TW

  const size_t frame = (size_t)blockIdx.z * w * h;
  const uint8_t* src = img + frame;
  const int x0 = blockIdx.x * B5_TW, y0 = blockIdx.y * B5_TH;
  const int tid = threadIdx.x;

  // ---- load input tile (32-bit words where the span is inside the image)
  const bool wordable = ((w & 3) == 0);
  for (int i = tid; i < B5_IH * (B5_IW / 4); i += B5_THREADS) {
    int r = i / (B5_IW / 4), c4 = i % (B5_IW / 4);
    int gy = refl101(y0 - 3 + r, h);
    int gx = x0 - 4 + 4 * c4;
    uint32_t v;
    if (wordable && gx >= 0 && gx + 3 < w) {
      v = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)gy * w + gx));
    } else {
      v = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) v |= (uint32_t)__ldg(src + (size_t)gy * w + refl101(gx + b, w)) << (8 * b);
    }
    *reinterpret_cast<uint32_t*>(&s_in[r][4 * c4]) = v;
  }
  __syncthreads();

  if (do_blur) {
    // ---- horizontal pass: columns x0-1 .. x0+TW  (s_in column = gx - (x0-4))
    for (int i = tid; i < B5_IH * (B5_TW + 2); i += B5_THREADS) {
      int r = i / (B5_TW + 2), c = i % (B5_TW + 2);
      const uint8_t* p = &s_in[r][c + 3 - 2];  // gx = x0-1+c -> col c+3
      s_hb[r][c] = (uint16_t)(14 * (p[0] + p[4]) + 62 * (p[1] + p[3]) + 104 * p[2]);
    }
    __syncthreads();
    // ---- vertical pass: rows y0-1 .. y0+TH
    for (int i = tid; i < (B5_TH + 2) * (B5_TW + 2); i += B5_THREADS) {
      int r = i / (B5_TW + 2), c = i % (B5_TW + 2);
      uint32_t s = 14u * (s_hb[r][c] + s_hb[r + 4][c]) + 62u * (s_hb[r + 1][c] + s_hb[r + 3][c]) + 104u * s_hb[r + 2][c];
      s_bl[r][c + 1] = (uint8_t)((s + 32768u) >> 16);
    }
  } else {
    for (int i = tid; i < (B5_TH + 2) * (B5_TW + 2); i += B5_THREADS) {
      int r = i / (B5_TW + 2), c = i % (B5_TW + 2);
      s_bl[r][c + 1] = s_in[r + 2][c + 3];
    }
  }
  __syncthreads();

  // ---- Sobel + stores: each thread handles 4 consecutive pixels of a row
  for (int i = tid; i < B5_TH * (B5_TW / 4); i += B5_THREADS) {
    int r = i >> 4, c = 4 * (i & 15);  // B5_TW / 4 == 16
    int gy = y0 + r, gx = x0 + c;
    if (gy >= h || gx >= w) continue;
    uint32_t packed = 0;
    short2 g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // blurred pixel (gx+k, gy) is s_bl[r+1][c+k+2]
      const uint8_t* a = &s_bl[r][c + k + 1];      // row above, x-1
      const uint8_t* m = &s_bl[r + 1][c + k + 1];  // same row
      const uint8_t* b = &s_bl[r + 2][c + k + 1];  // row below
      int dx = (a[2] + 2 * m[2] + b[2]) - (a[0] + 2 * m[0] + b[0]);
      int dy = (b[0] + 2 * b[1] + b[2]) - (a[0] + 2 * a[1] + a[2]);
      g[k] = make_short2((short)dx, (short)dy);
      packed |= (uint32_t)m[1] << (8 * k);
    }
    size_t o = frame + (size_t)gy * w + gx;
    if (wordable && gx + 3 < w) {
      *reinterpret_cast<uint32_t*>(pyr + o) = packed;
      *reinterpret_cast<uint4*>(grad + o) =
          make_uint4(*reinterpret_cast<uint32_t*>(&g[0]), *reinterpret_cast<uint32_t*>(&g[1]),
                     *reinterpret_cast<uint32_t*>(&g[2]), *reinterpret_cast<uint32_t*>(&g[3]));
    } else {
      for (int k = 0; k < 4 && gx + k < w; ++k) {
        pyr[o + k] = (uint8_t)(packed >> (8 * k));
        grad[o + k] = g[k];
      }
    }
  }
}

void launch_blur5_sobel(const uint8_t* img, uint8_t* pyr, short2* grad, int w, int h, int batch,
                        int do_blur, cudaStream_t st) {
  dim3 grid((w + B5_TW - 1) / B5_TW, (h + B5_TH - 1) / B5_TH, batch);
  blur5_sobel_kernel<<<grid, B5_THREADS, 0, st>>>(img, pyr, grad, w, h, do_blur);
}

// ---------------------------------------------------------------------------
// pyrDown to (w/2, h/2): [1 4 6 4 1]^2 centred on (2x,2y), REFLECT_101 on the
// source, (s+128)>>8.  Upper octaves only (1/4 of the pixels per level).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pyrdown_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int w, int h) {
  const int dw = w / 2, dh = h / 2;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dw || y >= dh) return;
  const uint8_t* s = src + (size_t)blockIdx.z * w * h;
  int xs[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) xs[i] = refl101(2 * x + i - 2, w);
  int acc = 0;
  const int k[5] = {1, 4, 6, 4, 1};
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const uint8_t* row = s + (size_t)refl101(2 * y + j - 2, h) * w;
    int rs = __ldg(row + xs[0]) + 4 * __ldg(row + xs[1]) + 6 * __ldg(row + xs[2]) + 4 * __ldg(row + xs[3]) +
             __ldg(row + xs[4]);
    acc += k[j] * rs;
  }
  dst[(size_t)blockIdx.z * dw * dh + (size_t)y * dw + x] = (uint8_t)((acc + 128) >> 8);
}

void launch_pyrdown(const uint8_t* src, uint8_t* dst, int w, int h, int batch, cudaStream_t st) {
  dim3 grid((w / 2 + 31) / 32, (h / 2 + 7) / 8, batch);
  pyrdown_kernel<<<grid, 256, 0, st>>>(src, dst, w, h);
}

// Sobel 3x3 -> short2 (dx,dy), REFLECT_101.  Upper octaves only.
__global__ void __launch_bounds__(256)
sobel_kernel(const uint8_t* __restrict__ src, short2* __restrict__ grad, int w, int h) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const uint8_t* s = src + (size_t)blockIdx.z * w * h;
  const int xm = refl101(x - 1, w), xp = refl101(x + 1, w);
  const uint8_t* ra = s + (size_t)refl101(y - 1, h) * w;
  const uint8_t* rm = s + (size_t)y * w;
  const uint8_t* rb = s + (size_t)refl101(y + 1, h) * w;
  int a0 = __ldg(ra + xm), a1 = __ldg(ra + x), a2 = __ldg(ra + xp);
  int m0 = __ldg(rm + xm), m2 = __ldg(rm + xp);
  int b0 = __ldg(rb + xm), b1 = __ldg(rb + x), b2 = __ldg(rb + xp);
  int dx = (a2 + 2 * m2 + b2) - (a0 + 2 * m0 + b0);
  int dy = (b0 + 2 * b1 + b2) - (a0 + 2 * a1 + a2);
  grad[(size_t)blockIdx.z * w * h + (size_t)y * w + x] = make_short2((short)dx, (short)dy);
}

void launch_sobel(const uint8_t* src, short2* grad, int w, int h, int batch, cudaStream_t st) {
  dim3 grid((w + 31) / 32, (h + 7) / 8, batch);
  sobel_kernel<<<grid, 256, 0, st>>>(src, grad, w, h);
}

// ---------------------------------------------------------------------------
// K2: LSD pre-scaling = GaussianBlur(7x7, sigma 0.75) (fixed-point taps
// [0 4 56 136 56 4 0]/256) fused with resize(0.8, INTER_LINEAR_EXACT): source
// coordinate f = 1.25 d + 0.125, weights in eighths, one round-half-up.
// A 64x16 destination tile reads an 81x21 block of blurred pixels (+2 halo).
// Algorithmic bytes: read P, write 0.64 P.
// ---------------------------------------------------------------------------
constexpr int SC_TW = 64, SC_TH = 16, SC_THREADS = 256;
constexpr int SC_GW = 81, SC_GH = 21;        // blurred block
constexpr int SC_IW = SC_GW + 4, SC_IH = SC_GH + 4;

__global__ void __launch_bounds__(SC_THREADS)
scale08_kernel(const uint8_t* __restrict__ src_, uint8_t* __restrict__ dst_, int w, int h, int ws, int hs) {
  __shared__ uint8_t s_in[SC_IH][SC_IW + 3];
  __shared__ uint16_t s_hb[SC_IH][SC_GW + 1];
  __shared__ uint8_t s_g[SC_GH][SC_GW + 3];
  const uint8_t* src = src_ + (size_t)blockIdx.z * w * h;
  uint8_t* dst = dst_ + (size_t)blockIdx.z * ws * hs;
  const int dx0 = blockIdx.x * SC_TW, dy0 = blockIdx.y * SC_TH;
  const int sx0 = (10 * dx0 + 1) >> 3, sy0 = (10 * dy0 + 1) >> 3;  // first source col/row of the tile
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;  // 32 x 8 threads
  for (int r = ty; r < SC_IH; r += 8) {
    const uint8_t* srow = src + (size_t)refl101(sy0 - 2 + r, h) * w;
    for (int c = tx; c < SC_IW; c += 32) s_in[r][c] = __ldg(srow + refl101(sx0 - 2 + c, w));
  }
  __syncthreads();
  for (int r = ty; r < SC_IH; r += 8)
    for (int c = tx; c < SC_GW; c += 32) {
      const uint8_t* p = &s_in[r][c];
      s_hb[r][c] = (uint16_t)(4 * (p[0] + p[4]) + 56 * (p[1] + p[3]) + 136 * p[2]);
    }
  __syncthreads();
  for (int r = ty; r < SC_GH; r += 8)
    for (int c = tx; c < SC_GW; c += 32) {
      uint32_t s = 4u * (s_hb[r][c] + s_hb[r + 4][c]) + 56u * (s_hb[r + 1][c] + s_hb[r + 3][c]) + 136u * s_hb[r + 2][c];
      s_g[r][c] = (uint8_t)((s + 32768u) >> 16);
    }
  __syncthreads();
  for (int i = tid; i < SC_TH * SC_TW; i += SC_THREADS) {
    int r = i >> 6, c = i & 63;  // SC_TW == 64
    int dx = dx0 + c, dy = dy0 + r;
    if (dx >= ws || dy >= hs) continue;
    int fx8 = 10 * dx + 1, fy8 = 10 * dy + 1;
    int ix = fx8 >> 3, ax = fx8 & 7, iy = fy8 >> 3, ay = fy8 & 7;
    if (ix >= w - 1) { ix = w - 1; ax = 0; }
    if (iy >= h - 1) { iy = h - 1; ay = 0; }
    int ix1 = min(ix + 1, w - 1), iy1 = min(iy + 1, h - 1);
    int lx = ix - sx0, lx1 = ix1 - sx0, ly = iy - sy0, ly1 = iy1 - sy0;
    int s = (8 - ay) * ((8 - ax) * s_g[ly][lx] + ax * s_g[ly][lx1]) +
            ay * ((8 - ax) * s_g[ly1][lx] + ax * s_g[ly1][lx1]);
    dst[(size_t)dy * ws + dx] = (uint8_t)((s + 32) >> 6);
  }
}

void launch_scale08(const uint8_t* src, uint8_t* dst, int w, int h, int ws, int hs, int batch,
                    cudaStream_t st) {
  dim3 grid((ws + SC_TW - 1) / SC_TW, (hs + SC_TH - 1) / SC_TH, batch);
  scale08_kernel<<<grid, SC_THREADS, 0, st>>>(src, dst, w, h, ws, hs);
}

}  // namespace vpl
