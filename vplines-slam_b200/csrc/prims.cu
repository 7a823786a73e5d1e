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
#include "vpl_tma.cuh"

namespace vpl {

// ---------------------------------------------------------------------------
// Separable 5-tap fixed-point Gaussian (taps W0 W1 W2 W1 W0, sum 256) of a byte tile held in
// shared memory, with packed arithmetic.  Both passes are exact integer sums and there is a
// single rounding at the end, so doing the vertical pass first gives the same bytes as
// OpenCV's row-then-column order.
//   vertical pass  : one 32-bit word = 4 pixels; even and odd bytes are split into two
//                    16-bit lanes each (SWAR): 14*(E0+E4) + 62*(E1+E3) + 104*E2 never exceeds
//                    255*256 = 65280 per lane, so there is no carry between lanes;
//   horizontal pass: the 16-bit column sums are consumed two at a time by dp2a (u16 x u8 dot
//                    product, 32-bit accumulate), four output pixels per thread, then
//                    (v + 32768) >> 16.
// s_in : in_rows x (4*in_words) bytes, pitch in_pitch bytes (multiple of 4)
// s_v  : (in_rows-4) x (4*in_words) u16, pitch v_pitch elements (multiple of 4)
// s_out: (in_rows-4) x 4*(in_words-1) bytes; s_out[r][c] is the blurred pixel of input
//        position (r+2, c+2); pitch out_pitch bytes (multiple of 4)
// ---------------------------------------------------------------------------
template <int W0, int W1, int W2>
__device__ __forceinline__ void blur5_tile_packed(const uint8_t* s_in, int in_pitch, int in_rows, int in_words,
                                                  uint16_t* s_v, int v_pitch, uint8_t* s_out, int out_pitch,
                                                  int tid, int nthreads) {
  const int vrows = in_rows - 4;
  for (int i = tid; i < vrows * in_words; i += nthreads) {
    const int r = i / in_words, wc = i - r * in_words;
    const uint8_t* p = s_in + r * in_pitch + 4 * wc;
    const uint32_t w0 = *reinterpret_cast<const uint32_t*>(p);
    const uint32_t w1 = *reinterpret_cast<const uint32_t*>(p + in_pitch);
    const uint32_t w2 = *reinterpret_cast<const uint32_t*>(p + 2 * in_pitch);
    const uint32_t w3 = *reinterpret_cast<const uint32_t*>(p + 3 * in_pitch);
    const uint32_t w4 = *reinterpret_cast<const uint32_t*>(p + 4 * in_pitch);
    const uint32_t M = 0x00ff00ffu;
    uint32_t e = (uint32_t)W0 * ((w0 & M) + (w4 & M)) + (uint32_t)W1 * ((w1 & M) + (w3 & M)) + (uint32_t)W2 * (w2 & M);
    uint32_t o = (uint32_t)W0 * (((w0 >> 8) & M) + ((w4 >> 8) & M)) + (uint32_t)W1 * (((w1 >> 8) & M) + ((w3 >> 8) & M)) +
                 (uint32_t)W2 * ((w2 >> 8) & M);
    // e = (px0, px2), o = (px1, px3) as 16-bit lanes -> natural order (px0,px1), (px2,px3)
    uint2 v;
    v.x = __byte_perm(e, o, 0x5410);
    v.y = __byte_perm(e, o, 0x7632);
    *reinterpret_cast<uint2*>(s_v + r * v_pitch + 4 * wc) = v;
  }
  __syncthreads();
  constexpr uint32_t B01 = (uint32_t)W0 | ((uint32_t)W1 << 8);  // (W0, W1)
  constexpr uint32_t B21 = (uint32_t)W2 | ((uint32_t)W1 << 8);  // (W2, W1)
  constexpr uint32_t B0_ = (uint32_t)W0;                         // (W0, 0)
  constexpr uint32_t B_0 = ((uint32_t)W0 << 8);                  // (0, W0)
  constexpr uint32_t B12 = (uint32_t)W1 | ((uint32_t)W2 << 8);  // (W1, W2)
  constexpr uint32_t B10 = (uint32_t)W1 | ((uint32_t)W0 << 8);  // (W1, W0)
  const int groups = in_words - 1;
  for (int i = tid; i < vrows * groups; i += nthreads) {
    const int r = i / groups, g = i - r * groups;
    const uint2 a = *reinterpret_cast<const uint2*>(s_v + r * v_pitch + 4 * g);      // v0..v3
    const uint2 c = *reinterpret_cast<const uint2*>(s_v + r * v_pitch + 4 * g + 4);  // v4..v7
    uint32_t o0 = __dp2a_lo(a.x, B01, __dp2a_lo(a.y, B21, __dp2a_lo(c.x, B0_, 32768u)));
    uint32_t o1 = __dp2a_lo(a.x, B_0, __dp2a_lo(a.y, B12, __dp2a_lo(c.x, B10, 32768u)));
    uint32_t o2 = __dp2a_lo(a.y, B01, __dp2a_lo(c.x, B21, __dp2a_lo(c.y, B0_, 32768u)));
    uint32_t o3 = __dp2a_lo(a.y, B_0, __dp2a_lo(c.x, B12, __dp2a_lo(c.y, B10, 32768u)));
    // bits 16..23 of each sum are the result bytes
    uint32_t lo = __byte_perm(o0, o1, 0x0062);  // byte2(o0), byte2(o1)
    uint32_t hi = __byte_perm(o2, o3, 0x0062);
    *reinterpret_cast<uint32_t*>(s_out + r * out_pitch + 4 * g) = __byte_perm(lo, hi, 0x5410);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// K1: GaussianBlur(5x5, sigma 1) fused with Sobel 3x3 (dx,dy int16).
// Fixed-point kernel [14 62 104 62 14]/256, (v + 32768) >> 16.  BORDER_REFLECT_101 on both
// stages; the reflect-extended input is symmetric about the border pixel and the kernel is
// symmetric, so blurring the extended tile yields exactly the reflected blurred halo the
// Sobel stage needs.  64x32 output tile, input tile 72x38 bytes loaded as 32-bit words.
// Algorithmic bytes: read P, write P (pyr) + 4P (grad) = 6P per frame.
// ---------------------------------------------------------------------------
constexpr int B5_TW = 64, B5_TH = 32, B5_THREADS = 256;
constexpr int B5_IW = B5_TW + 8;  // input tile: x0-4 .. x0+TW+3 (word aligned), 18 words
constexpr int B5_IH = B5_TH + 6;  // y0-3 .. y0+TH+2
constexpr int B5_OW = B5_IW - 4;  // blurred tile: 68 columns, column c <-> gx = x0-2+c
constexpr int B5_OH = B5_IH - 4;  // 34 rows, row r <-> gy = y0-1+r

// Sobel 3x3 of the blurred tile + the stores: each thread handles 4 consecutive pixels of a row.
__device__ __forceinline__ void blur5_sobel_store(const uint8_t (*s_bl)[B5_IW], uint8_t* __restrict__ pyr,
                                                  short2* __restrict__ grad, size_t frame, int x0, int y0, int w, int h,
                                                  int tid) {
  const bool wordable = ((w & 3) == 0);
  // output pixel (x0+c+k, y0+r) is blurred column c+k+2, row r+1 of s_bl
  for (int i = tid; i < B5_TH * (B5_TW / 4); i += B5_THREADS) {
    int r = i >> 4, c = 4 * (i & 15);  // B5_TW / 4 == 16
    int gy = y0 + r, gx = x0 + c;
    if (gy >= h || gx >= w) continue;
    // bytes c .. c+7 of the three rows; the window needed is columns c+1 .. c+6
    uint2 ra, rm, rb;  // (c is a multiple of 4 only: two 32-bit loads per row)
    ra.x = *reinterpret_cast<const uint32_t*>(&s_bl[r][c]);     ra.y = *reinterpret_cast<const uint32_t*>(&s_bl[r][c + 4]);
    rm.x = *reinterpret_cast<const uint32_t*>(&s_bl[r + 1][c]); rm.y = *reinterpret_cast<const uint32_t*>(&s_bl[r + 1][c + 4]);
    rb.x = *reinterpret_cast<const uint32_t*>(&s_bl[r + 2][c]); rb.y = *reinterpret_cast<const uint32_t*>(&s_bl[r + 2][c + 4]);
    int C[6], D[6];  // column sums a+2m+b and column differences b-a for columns c+1 .. c+6
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int bi = j + 1;
      const int a = (bi < 4) ? (int)((ra.x >> (8 * bi)) & 0xff) : (int)((ra.y >> (8 * (bi - 4))) & 0xff);
      const int m = (bi < 4) ? (int)((rm.x >> (8 * bi)) & 0xff) : (int)((rm.y >> (8 * (bi - 4))) & 0xff);
      const int b = (bi < 4) ? (int)((rb.x >> (8 * bi)) & 0xff) : (int)((rb.y >> (8 * (bi - 4))) & 0xff);
      C[j] = a + 2 * m + b;
      D[j] = b - a;
    }
    short2 g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = make_short2((short)(C[k + 2] - C[k]), (short)(D[k] + 2 * D[k + 1] + D[k + 2]));
    const uint32_t packed = __byte_perm(rm.x, rm.y, 0x5432);  // blurred centre pixels c+2 .. c+5
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

__global__ void __launch_bounds__(B5_THREADS)
blur5_sobel_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ pyr,
                   short2* __restrict__ grad, int w, int h, int do_blur) {
  __shared__ __align__(16) uint8_t s_in[B5_IH][B5_IW];
  __shared__ __align__(16) uint16_t s_v[B5_OH][B5_IW];
  __shared__ __align__(16) uint8_t s_bl[B5_OH][B5_IW];

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
    blur5_tile_packed<14, 62, 104>(&s_in[0][0], B5_IW, B5_IH, B5_IW / 4, &s_v[0][0], B5_IW, &s_bl[0][0], B5_IW, tid,
                                   B5_THREADS);
  } else {
    for (int i = tid; i < B5_OH * (B5_OW / 4); i += B5_THREADS) {
      int r = i / (B5_OW / 4), c = 4 * (i % (B5_OW / 4));
      // s_bl[r][c] = s_in[r+2][c+2]: unaligned by two bytes
      uint32_t lo = *reinterpret_cast<const uint32_t*>(&s_in[r + 2][c]);
      uint32_t hi = *reinterpret_cast<const uint32_t*>(&s_in[r + 2][c + 4]);
      *reinterpret_cast<uint32_t*>(&s_bl[r][c]) = __byte_perm(lo, hi, 0x5432);
    }
    __syncthreads();
  }

  blur5_sobel_store(s_bl, pyr, grad, frame, x0, y0, w, h, tid);
}

// ---------------------------------------------------------------------------
// The same kernel with the input tile brought in by the TMA unit: one cp.async.bulk.tensor.3d per CTA moves the
// 96 x 38-byte box (x0-16 .. x0+79, y0-3 .. y0+34) of frame z into shared memory and signals an mbarrier; elements
// outside the image arrive as zeros and are then overwritten with their BORDER_REFLECT_101 mirror images, which lie in
// the same tile (the halo is 4 pixels, the margins 12 and more).  No per-thread address arithmetic, no refl101 per
// byte.  The unit wants the box to start on a 16-byte boundary of the row (measured on B200: any other inner
// coordinate raises "illegal instruction", tools/probe/), hence x0-16 and a width of 96 for the 72 bytes used, and
// a row pitch that is a multiple of 16 bytes (w % 16 == 0), otherwise the kernel above runs.
// ---------------------------------------------------------------------------
constexpr int B5T_PITCH = 96;  // box width in bytes (multiple of 16)
constexpr int B5T_X = 16;      // tile column c <-> gx = x0 - B5T_X + c; the columns used are B5T_X-4 .. B5T_X+67

__global__ void __launch_bounds__(B5_THREADS)
blur5_sobel_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t* __restrict__ pyr, short2* __restrict__ grad,
                       int w, int h, int do_blur) {
  __shared__ __align__(128) uint8_t s_in[B5_IH][B5T_PITCH];
  __shared__ __align__(16) uint16_t s_v[B5_OH][B5_IW];
  __shared__ __align__(16) uint8_t s_bl[B5_OH][B5_IW];
  __shared__ __align__(8) unsigned long long s_bar;

  const size_t frame = (size_t)blockIdx.z * w * h;
  const int x0 = blockIdx.x * B5_TW, y0 = blockIdx.y * B5_TH;
  const int tid = threadIdx.x;
  tma_load_box_3d(&tmap, &s_in[0][0], &s_bar, x0 - B5T_X, y0 - 3, (int)blockIdx.z, B5_IH * B5T_PITCH);
  if (x0 - 4 < 0 || x0 + B5_TW + 4 > w || y0 - 3 < 0 || y0 + B5_TH + 3 > h)
    tma_reflect_fix(&s_in[0][0], B5T_PITCH, B5_IH, B5T_X - 4, B5_IW, x0 - B5T_X, y0 - 3, w, h, B5_THREADS);

  if (do_blur) {
    blur5_tile_packed<14, 62, 104>(&s_in[0][B5T_X - 4], B5T_PITCH, B5_IH, B5_IW / 4, &s_v[0][0], B5_IW, &s_bl[0][0], B5_IW,
                                   tid, B5_THREADS);
  } else {
    for (int i = tid; i < B5_OH * (B5_OW / 4); i += B5_THREADS) {
      int r = i / (B5_OW / 4), c = 4 * (i % (B5_OW / 4));
      uint32_t lo = *reinterpret_cast<const uint32_t*>(&s_in[r + 2][c + B5T_X - 4]);
      uint32_t hi = *reinterpret_cast<const uint32_t*>(&s_in[r + 2][c + B5T_X]);
      *reinterpret_cast<uint32_t*>(&s_bl[r][c]) = __byte_perm(lo, hi, 0x5432);
    }
    __syncthreads();
  }
  blur5_sobel_store(s_bl, pyr, grad, frame, x0, y0, w, h, tid);
}

void launch_blur5_sobel(const uint8_t* img, uint8_t* pyr, short2* grad, int w, int h, int batch,
                        int do_blur, cudaStream_t st) {
  dim3 grid((w + B5_TW - 1) / B5_TW, (h + B5_TH - 1) / B5_TH, batch);
  CUtensorMap tm;
  if (make_u8_frames_tmap(&tm, img, w, h, batch, B5T_PITCH, B5_IH))
    blur5_sobel_tma_kernel<<<grid, B5_THREADS, 0, st>>>(tm, pyr, grad, w, h, do_blur);
  else
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
constexpr int SC_GH = 21;                // blurred block: columns sx0-2 .. sx0+81 (81 needed), rows sy0 .. sy0+20
constexpr int SC_IW = 88, SC_IH = 25;   // input block: columns sx0-4 .. sx0+83 (22 words), rows sy0-2 .. sy0+22

// blur of the staged block + the 0.8 resampling of one 64 x 16 destination tile; s_in(r, c) = source pixel
// (sx0 - 4 + c, sy0 - 2 + r), row pitch in_pitch bytes
__device__ __forceinline__ void scale08_tile(const uint8_t* s_in, int in_pitch, uint16_t (*s_v)[88], uint8_t (*s_g)[88],
                                             uint8_t* __restrict__ dst, int dx0, int dy0, int sx0, int sy0, int w, int h,
                                             int ws, int hs, int tid);

__global__ void __launch_bounds__(SC_THREADS)
scale08_kernel(const uint8_t* __restrict__ src_, uint8_t* __restrict__ dst_, int w, int h, int ws, int hs) {
  __shared__ __align__(16) uint8_t s_in[SC_IH][SC_IW];
  __shared__ __align__(16) uint16_t s_v[SC_GH][SC_IW];
  __shared__ __align__(16) uint8_t s_g[SC_GH][SC_IW];  // s_g[r][c]: blurred pixel (sx0 - 2 + c, sy0 + r)
  const uint8_t* src = src_ + (size_t)blockIdx.z * w * h;
  uint8_t* dst = dst_ + (size_t)blockIdx.z * ws * hs;
  const int dx0 = blockIdx.x * SC_TW, dy0 = blockIdx.y * SC_TH;
  const int sx0 = (10 * dx0 + 1) >> 3, sy0 = (10 * dy0 + 1) >> 3;  // first source col/row of the tile (sx0 % 4 == 0)
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;  // 32 x 8 threads
  const bool wordable = ((w & 3) == 0);
  for (int r = ty; r < SC_IH; r += 8) {
    const uint8_t* srow = src + (size_t)refl101(sy0 - 2 + r, h) * w;
    if (tx < SC_IW / 4) {
      const int gx = sx0 - 4 + 4 * tx;
      uint32_t v;
      if (wordable && gx >= 0 && gx + 3 < w) {
        v = __ldg(reinterpret_cast<const uint32_t*>(srow + gx));
      } else {
        v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) v |= (uint32_t)__ldg(srow + refl101(gx + b, w)) << (8 * b);
      }
      *reinterpret_cast<uint32_t*>(&s_in[r][4 * tx]) = v;
    }
  }
  __syncthreads();
  scale08_tile(&s_in[0][0], SC_IW, s_v, s_g, dst, dx0, dy0, sx0, sy0, w, h, ws, hs, tid);
}

__device__ __forceinline__ void scale08_tile(const uint8_t* s_in, int in_pitch, uint16_t (*s_v)[88], uint8_t (*s_g)[88],
                                             uint8_t* __restrict__ dst, int dx0, int dy0, int sx0, int sy0, int w, int h,
                                             int ws, int hs, int tid) {
  blur5_tile_packed<4, 56, 136>(s_in, in_pitch, SC_IH, SC_IW / 4, &s_v[0][0], SC_IW, &s_g[0][0], SC_IW, tid, SC_THREADS);
  for (int i = tid; i < SC_TH * SC_TW; i += SC_THREADS) {
    int r = i >> 6, c = i & 63;  // SC_TW == 64
    int dx = dx0 + c, dy = dy0 + r;
    if (dx >= ws || dy >= hs) continue;
    int fx8 = 10 * dx + 1, fy8 = 10 * dy + 1;
    int ix = fx8 >> 3, ax = fx8 & 7, iy = fy8 >> 3, ay = fy8 & 7;
    if (ix >= w - 1) { ix = w - 1; ax = 0; }
    if (iy >= h - 1) { iy = h - 1; ay = 0; }
    int ix1 = min(ix + 1, w - 1), iy1 = min(iy + 1, h - 1);
    int lx = ix - sx0 + 2, lx1 = ix1 - sx0 + 2, ly = iy - sy0, ly1 = iy1 - sy0;
    int s = (8 - ay) * ((8 - ax) * s_g[ly][lx] + ax * s_g[ly][lx1]) +
            ay * ((8 - ax) * s_g[ly1][lx] + ax * s_g[ly1][lx1]);
    dst[(size_t)dy * ws + dx] = (uint8_t)((s + 32) >> 6);
  }
}

// The same with the source block brought in by the TMA unit: sx0 is a multiple of 80 (dx0 of 64), so the box starts
// at sx0 - 16 and is 112 bytes wide for the 88 used (columns 12 .. 99).
constexpr int SCT_PITCH = 112, SCT_X = 16;
__global__ void __launch_bounds__(SC_THREADS)
scale08_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t* __restrict__ dst_, int w, int h, int ws, int hs) {
  __shared__ __align__(128) uint8_t s_in[SC_IH][SCT_PITCH];
  __shared__ __align__(16) uint16_t s_v[SC_GH][SC_IW];
  __shared__ __align__(16) uint8_t s_g[SC_GH][SC_IW];
  __shared__ __align__(8) unsigned long long s_bar;
  uint8_t* dst = dst_ + (size_t)blockIdx.z * ws * hs;
  const int dx0 = blockIdx.x * SC_TW, dy0 = blockIdx.y * SC_TH;
  const int sx0 = (10 * dx0 + 1) >> 3, sy0 = (10 * dy0 + 1) >> 3;
  tma_load_box_3d(&tmap, &s_in[0][0], &s_bar, sx0 - SCT_X, sy0 - 2, (int)blockIdx.z, SC_IH * SCT_PITCH);
  if (sx0 - 4 < 0 || sx0 + SC_IW - 4 > w || sy0 - 2 < 0 || sy0 - 2 + SC_IH > h)
    tma_reflect_fix(&s_in[0][0], SCT_PITCH, SC_IH, SCT_X - 4, SC_IW, sx0 - SCT_X, sy0 - 2, w, h, SC_THREADS);
  scale08_tile(&s_in[0][SCT_X - 4], SCT_PITCH, s_v, s_g, dst, dx0, dy0, sx0, sy0, w, h, ws, hs, threadIdx.x);
}

void launch_scale08(const uint8_t* src, uint8_t* dst, int w, int h, int ws, int hs, int batch,
                    cudaStream_t st) {
  dim3 grid((ws + SC_TW - 1) / SC_TW, (hs + SC_TH - 1) / SC_TH, batch);
  CUtensorMap tm;
  if (make_u8_frames_tmap(&tm, src, w, h, batch, SCT_PITCH, SC_IH))
    scale08_tma_kernel<<<grid, SC_THREADS, 0, st>>>(tm, dst, w, h, ws, hs);
  else
    scale08_kernel<<<grid, SC_THREADS, 0, st>>>(src, dst, w, h, ws, hs);
}

}  // namespace vpl
