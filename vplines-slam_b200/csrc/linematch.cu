// linematch.cu -- the reference's real line matcher on the device (sm_100a).
//
// Replaces LineMatching::Matching (/root/reference/line_matching/src/line_matching.cpp:605-690,
// cited as lm.cpp), KLT::calc2D (klt.cpp:491-628) and LKTrackerInvoker2D::operator()
// (lk_tracker_invoker_2d.cpp:28-480, cited as lk2d.cpp), which the tracker reaches through
// match_line_match (feature_tracker/src/line_feature_tracker.cpp:115, :291-313) with
// illumination_adapt = true, topological_filter = true and no affine models.
// CPU restatement: oracle/orc_linematch.c (pinned bit for bit against the reference's own code).
//
// A "pair" p matches the lines of frame ref(p) = p*pstride to the lines of frame cur(p) = ref(p)+1
// (pstride 1: consecutive frames of a sequence; 2: independent pairs).  Stages, each one launch
// over the whole batch:
//   klt_level0 / klt_pyrdown / klt_scharr   cv::buildOpticalFlowPyramid (13-px REFLECT_101 border
//                     stored physically, so that windows never test bounds) and the Scharr
//                     derivative with its zero border (klt.cpp:42-122, :613)              HBM-bound
//   lm_anchor_kernel  anchors every `step` px along each reference line (lm.cpp:531-599)
//   klt_track_kernel  per pyramid level, ONE THREAD PER ANCHOR: the reference's float sums over
//                     the 13x13 window are sequential (y, x) single-precision accumulations, so
//                     the window loop stays in one thread and the 32 lanes of a warp carry 32
//                     anchors; persistent warps refill idle lanes from a queue over all pairs (the
//                     iteration count varies from 2 to 30); the windows live in shared memory ([element][lane], 43 KB per warp);
//                     each source byte of a window is loaded once (row-blended bilinear
//                     interpolation in exact integers)
//   lm_vote_kernel    one CTA per pair: closest current line per tracked anchor (lm.cpp:48-86),
//                     vote per reference line (:88-133), topological filter (:267-410, :656-665)
// Bit-exact vs the oracle: integer window arithmetic, float/double sequences in the reference's
// order with -fmad=false, IEEE sqrt/div.
#include <float.h>
#include <limits.h>
#include <stdlib.h>

#include "vpl_common.cuh"

namespace vpl {

namespace {

constexpr int WIN = 13;           // LineMatching::Matching sets 13x13 (lm.cpp:631)
constexpr int NWIN = WIN * WIN;
constexpr int W_BITS = 14;
#define LM_DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))  // CV_DESCALE, klt.h:39

// two unsigned 16-bit integers (lo | hi << 16) -> two floats, minus `bias`, without the conversion
// unit: the 16 bits are spliced under the exponent of 1.5 * 2^23 (exact), then ONE packed add
// subtracts 1.5 * 2^23 + bias from both (exact: every value involved is an integer below 2^24).
__device__ __forceinline__ void u16x2_to_float(unsigned packed, float neg_magic_bias, float& lo, float& hi) {
  const unsigned a = __byte_perm(packed, 0x4B400000u, 0x7610), b = __byte_perm(packed, 0x4B400000u, 0x7632);
  unsigned long long in, nb, out;
  const unsigned nbits = __float_as_uint(neg_magic_bias);
  asm("mov.b64 %0, {%1, %2};" : "=l"(in) : "r"(a), "r"(b));
  asm("mov.b64 %0, {%1, %1};" : "=l"(nb) : "r"(nbits));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(out) : "l"(in), "l"(nb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(out));
}

__device__ __forceinline__ int cv_floor_d(float v) {  // SSE cvFloor: INT_MIN on NaN / overflow
  if (!(v > -2147483648.0f && v < 2147483648.0f)) return INT_MIN;
  return (int)floorf(v);
}

// ---- pyramid ---------------------------------------------------------------------------------
// Level buffers: (h + 2 pad) rows of `stride` bytes (a multiple of 16); pixel (x, y) of the level is
// at row y + pad, column x + padx (padx = 16), so that rows and pixel 0 are 16-byte aligned and the
// kernels below move 4 pixels per thread.  Columns -pad .. w + pad - 1 hold valid border values.
__global__ void klt_level0_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ pyr, KltGeom G, int w, int h) {
  const int f = blockIdx.z;
  const int cw = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;  // word cw = columns 4cw .. 4cw+3
  const int st = G.stride[0];
  if (4 * cw >= st) return;
  const int sy = refl101(y - G.pad, h), x = 4 * cw - G.padx;
  const uint8_t* row = img + (size_t)f * w * h + (size_t)sy * w;
  unsigned v;
  if ((w & 3) == 0 && x >= 0 && x + 3 < w) {
    v = *reinterpret_cast<const unsigned*>(row + x);
  } else {
    v = (unsigned)row[refl101(x, w)] | ((unsigned)row[refl101(x + 1, w)] << 8) | ((unsigned)row[refl101(x + 2, w)] << 16) |
        ((unsigned)row[refl101(x + 3, w)] << 24);
  }
  *reinterpret_cast<unsigned*>(pyr + (size_t)f * G.img_frame + G.img_off[0] + (size_t)y * st + 4 * cw) = v;
}

// level l (padded) from level l-1 (padded): cv::pyrDown to ((w+1)/2, (h+1)/2), then the border
__global__ void klt_pyrdown_kernel(uint8_t* __restrict__ pyr, KltGeom G, int l) {
  const int f = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int st = G.stride[l];
  if (x >= st) return;
  const int dx = refl101(x - G.padx, G.w[l]), dy = refl101(y - G.pad, G.h[l]);
  const int sst = G.stride[l - 1];
  const uint8_t* src = pyr + (size_t)f * G.img_frame + G.img_off[l - 1];
  // source rows/cols 2d-2 .. 2d+2 lie inside the stored REFLECT_101 border of level l-1, also
  // beyond its far edge when the size is odd (2d+2 can reach w+1 <= w+pad-1)
  int s = 0;
#pragma unroll
  for (int j = -2; j <= 2; j++) {
    const uint8_t* row = src + (size_t)(2 * dy + j + G.pad) * sst + (2 * dx + G.padx);
    const int kj = j == 0 ? 6 : (j == -1 || j == 1) ? 4 : 1;
    s += kj * (row[-2] + 4 * row[-1] + 6 * row[0] + 4 * row[1] + row[2]);
  }
  pyr[(size_t)f * G.img_frame + G.img_off[l] + (size_t)y * st + x] = (uint8_t)((s + 128) >> 8);
}

// KLT::calcSharrDeriv + copyMakeBorder(BORDER_CONSTANT): zero outside the level.  4 pixels per
// thread: 3 rows x 3 aligned words in, one 128-bit store out.
__global__ void klt_scharr_kernel(const uint8_t* __restrict__ pyr, short2* __restrict__ deriv, KltGeom G, int l) {
  const int f = blockIdx.z;
  const int cw = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int st = G.stride[l];
  if (4 * cw >= st) return;
  const int iy = y - G.pad;
  int4 o = make_int4(0, 0, 0, 0);
  const int xb = 4 * cw - G.padx;  // level x of the first of the 4 pixels
  if (iy >= 0 && iy < G.h[l] && xb + 3 >= 0 && xb < G.w[l]) {
    const uint8_t* p = pyr + (size_t)f * G.img_frame + G.img_off[l] + (size_t)y * st + 4 * cw;
    // bytes x-4 .. x+7 of the three rows (cw >= 1 whenever a pixel of the word is inside the level)
    unsigned r0[3], r1[3], r2[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      r0[k] = *reinterpret_cast<const unsigned*>(p - st + 4 * (k - 1));
      r1[k] = *reinterpret_cast<const unsigned*>(p + 4 * (k - 1));
      r2[k] = *reinterpret_cast<const unsigned*>(p + st + 4 * (k - 1));
    }
    int res[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      // byte j of the 12-byte span = column x-4+j; pixel k sits at j = 4 + k
      auto at = [&](const unsigned* r, int j) { return (int)((r[j >> 2] >> (8 * (j & 3))) & 0xffu); };
      const int j = 4 + k;
      const int a = at(r0, j - 1), b = at(r0, j), c = at(r0, j + 1), d = at(r1, j - 1), e = at(r1, j + 1);
      const int g = at(r2, j - 1), hh = at(r2, j), i = at(r2, j + 1);
      short2 v = make_short2(0, 0);
      if (xb + k >= 0 && xb + k < G.w[l]) {
        // dIx = [3 10 3]^T (rows) x [-1 0 1];  dIy = [-1 0 1]^T x [3 10 3]
        v.x = (short)(((c + i) * 3 + e * 10) - ((a + g) * 3 + d * 10));
        v.y = (short)(((g - a) + (i - c)) * 3 + (hh - b) * 10);
      }
      res[k] = *reinterpret_cast<int*>(&v);
    }
    o = make_int4(res[0], res[1], res[2], res[3]);
  }
  *reinterpret_cast<int4*>(deriv + (size_t)f * G.deriv_frame + G.deriv_off[l] + (size_t)y * st + 4 * cw) = o;
}

// ---- anchors, lm.cpp:531-599 --------------------------------------------------------------------
__global__ void __launch_bounds__(256) lm_anchor_kernel(const VplLine* __restrict__ lines, const int* __restrict__ counts,
                                                        int cap, LmBuffers B, LmParams P, int pstride) {
  const int p = blockIdx.x;
  const int fr = p * pstride;
  const int n = min(counts[fr], cap);
  const VplLine* L = lines + (size_t)fr * cap;
  int* kp_start = B.kp_start + (size_t)p * (cap + 1);
  // exclusive scan of the per-line anchor counts (iter_num + 2), chunks of 1024 lines
  int base = 0;
  for (int c0 = 0; c0 < n; c0 += 1024) {
    int i = c0 + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (i + k < n) ? (int)(L[i + k].length / P.step) + 2 : 0;
    int t = v[0] + v[1] + v[2] + v[3];
    // block scan over 256 partial sums
    __shared__ int s_part[256];
    s_part[threadIdx.x] = t;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
      int a = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
      __syncthreads();
      s_part[threadIdx.x] += a;
      __syncthreads();
    }
    int ex = base + s_part[threadIdx.x] - t;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i + k < n) kp_start[i + k] = ex;
      ex += v[k];
    }
    base += s_part[255];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    kp_start[n] = base;
    B.n_kp[p] = base <= B.cap_kp ? base : 0;
    if (base > B.cap_kp) atomicExch(B.overflow, 1);
  }
  __syncthreads();
  if (base > B.cap_kp) return;
  float2* kps = B.kps + (size_t)p * B.cap_kp;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const VplLine l = L[i];
    float x1 = l.endpoint[0], y1 = l.endpoint[1], x2 = l.endpoint[2], y2 = l.endpoint[3];
    float px = x1, py = y1, len = l.length;
    float dirx = (x2 - x1) / len, diry = (y2 - y1) / len;
    float ddx = P.step * dirx, ddy = P.step * diry;
    int iter = (int)(len / P.step);
    int o = kp_start[i];
    for (int j = 0; j <= iter; j++) {
      kps[o++] = make_float2(px, py);
      px += ddx; py += ddy;
    }
    kps[o] = make_float2(x2, y2);
  }
}

// ---- per-point tracker ---------------------------------------------------------------------------
struct Wts {
  int w00, w01, w10, w11;
};
__device__ __forceinline__ Wts lk_weights(float a, float b) {  // lk2d.cpp:109-112
  Wts w;
  w.w00 = __float2int_rn((1.f - a) * (1.f - b) * (1 << W_BITS));
  w.w01 = __float2int_rn(a * (1.f - b) * (1 << W_BITS));
  w.w10 = __float2int_rn((1.f - a) * b * (1 << W_BITS));
  w.w11 = (1 << W_BITS) - w.w00 - w.w01 - w.w10;
  return w;
}

// bilinear 13x13 window of a padded u8 level at integer corner (ix, iy): every source byte is read
// once; the sums are exact integers, so blending rows first changes nothing (lk2d.cpp:338-348).
// Also returns the window's integer sum and sum of squares (for cv::meanStdDev).
// Window element t of lane l lives in the 32-bit word win[t * 32 + l] (shared memory, conflict-free):
// reference patch in the low half, current patch in the high half; `out` points at the lane's half.
__device__ __forceinline__ void lk_sample_u8(const uint8_t* __restrict__ img, int st, int ix, int iy, const Wts& w,
                                             unsigned short* __restrict__ out, int& sum, long long& sq) {
  int top[WIN];
  sum = 0;
  sq = 0;
  const uint8_t* row = img + (ptrdiff_t)iy * st + ix;
  // software pipeline: the bytes of row r+1 are in flight while row r is blended
  unsigned char nxt[WIN + 1];
#pragma unroll
  for (int x = 0; x <= WIN; x++) nxt[x] = row[x];
#pragma unroll 1
  for (int r = 0; r <= WIN; r++) {
    unsigned char cur_row[WIN + 1];
#pragma unroll
    for (int x = 0; x <= WIN; x++) cur_row[x] = nxt[x];
    row += st;
    if (r < WIN) {
#pragma unroll
      for (int x = 0; x <= WIN; x++) nxt[x] = row[x];
    }
    int prev = cur_row[0];
    unsigned rq = 0;  // 13 * 8160^2 < 2^31
#pragma unroll
    for (int x = 0; x < WIN; x++) {
      int cur = cur_row[x + 1];
      int t = prev * w.w00 + cur * w.w01;   // this row as the upper row of window row r
      int b = prev * w.w10 + cur * w.w11;   // ... and as the lower row of window row r-1
      if (r > 0) {
        int v = LM_DESCALE(top[x] + b, W_BITS - 5);
        out[((r - 1) * WIN + x) * 64] = (unsigned short)v;
        sum += v;
        rq += (unsigned)(v * v);
      }
      top[x] = t;
      prev = cur;
    }
    sq += rq;
  }
}

__device__ __forceinline__ void lk_norm_params(int si, long long qi, int sj, long long qj, float& alpha, float& beta) {
  // getImageNormParams (klt.cpp:4-10) over cv::meanStdDev of two CV_16S windows
  const double scale = 1.0 / NWIN;
  double mi = si * scale, mj = sj * scale;
  double vi = (double)qi * scale - mi * mi, vj = (double)qj * scale - mj * mj;
  double sdi = sqrt(vi > 0 ? vi : 0), sdj = sqrt(vj > 0 ? vj : 0);
  alpha = (float)(sdi / sdj);
  beta = (float)(mi - alpha * mj);
}

// exclusive prefix of the per-pair anchor counts: the work queue of the tracker runs over all pairs
__global__ void __launch_bounds__(1024) lm_pair_offsets_kernel(const int* __restrict__ n_kp, int n_pairs,
                                                               int* __restrict__ pair_off) {
  __shared__ int s_part[1024];
  const int per = (n_pairs + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(lo + per, n_pairs);
  int t = 0;
  for (int i = lo; i < hi; i++) t += n_kp[i];
  s_part[threadIdx.x] = t;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    int v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  int ex = s_part[threadIdx.x] - t;
  for (int i = lo; i < hi; i++) {
    pair_off[i] = ex;
    ex += n_kp[i];
  }
  if (threadIdx.x == 1023) pair_off[n_pairs] = s_part[1023];
}

// Persistent warps, one per CTA; the three 13x13 windows of a lane's anchor (reference patch, its
// Scharr pair, current patch: 1352 B) live in shared memory, 43 KB per CTA, 5 CTAs per SM.
//
// The number of LK iterations varies a lot between anchors (mean 4-9, but 4-7 % run all 30), so a
// warp that takes 32 anchors and waits for the slowest spends two thirds of its issue slots on
// idle lanes.  Instead every lane is a small state machine (idle / iterating / final error) and
// the warp pulls new anchors from a queue over all pairs of the batch whenever kRefill (16) lanes
// are idle (the set-up of a new anchor runs divergently, so it should not run for too few lanes); one trip of the loop = one window sampling + one accumulation pass for every busy lane,
// whatever its phase.  Per-anchor arithmetic is untouched, so results do not depend on the schedule.
constexpr int kTrackSmem = NWIN * 32 * (2 + 2 + 4);
__global__ void __launch_bounds__(32) klt_track_kernel(const uint8_t* __restrict__ pyr, const short2* __restrict__ deriv,
                                                       KltGeom G, LmBuffers B, LmParams P, int level, int pstride,
                                                       int n_pairs, int kRefill) {
  extern __shared__ __align__(16) unsigned char s_win[];
  const int lane = threadIdx.x;
  unsigned* IJw = reinterpret_cast<unsigned*>(s_win) + lane;          // {I, J} as two u16 (values 0..8160)
  unsigned short* Iw = reinterpret_cast<unsigned short*>(IJw);
  unsigned short* Jw = Iw + 1;
  unsigned* dIw = reinterpret_cast<unsigned*>(s_win + NWIN * 32 * 4) + lane;  // {dIx, dIy} + 32768 as two u16
  const int total = B.pair_off[n_pairs];
  int* queue = B.queue + level;
  const int st = G.stride[level], lw = G.w[level], lh = G.h[level];
  const float half = (WIN - 1) * 0.5f;
  const float FLT_SCALE = 1.f / (1 << 20);
  const float scale = (float)(1. / (1 << level));

  int phase = 0;  // 0 idle, 1 iterating, 2 final error (level 0 only)
  size_t k = 0;
  const uint8_t* J = nullptr;
  float nx = 0, ny = 0, pdx = 0, pdy = 0, outx = 0, outy = 0, A11 = 0, A12 = 0, A22 = 0, D = 0;
  int j = 0, sI = 0;
  long long qI = 0;
  bool exhausted = false;

  for (;;) {
    const unsigned idle = __ballot_sync(0xffffffffu, phase == 0);
    if (!exhausted && (__popc(idle) >= kRefill || idle == 0xffffffffu)) {
      const int cnt = __popc(idle);
      int base = 0;
      if (lane == 0) base = atomicAdd(queue, cnt);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + cnt >= total) exhausted = true;
      const int t = base + __popc(idle & ((1u << lane) - 1u));
      if (phase == 0 && t < total) {
        // ---- set-up of anchor t at this level, lk2d.cpp:45-300 ----
        int lo = 0, hi = n_pairs;  // largest p with pair_off[p] <= t
        while (hi - lo > 1) {
          int mid = (lo + hi) >> 1;
          if (B.pair_off[mid] <= t) lo = mid; else hi = mid;
        }
        const int p = lo, i = t - B.pair_off[lo];
        k = (size_t)p * B.cap_kp + i;
        const int fr = p * pstride, fc = fr + 1;
        // pointers to pixel (0,0) of the level inside its padded buffer
        const uint8_t* I = pyr + (size_t)fr * G.img_frame + G.img_off[level] + (size_t)G.pad * st + G.padx;
        J = pyr + (size_t)fc * G.img_frame + G.img_off[level] + (size_t)G.pad * st + G.padx;
        const short2* dI = deriv + (size_t)fr * G.deriv_frame + G.deriv_off[level] + (size_t)G.pad * st + G.padx;
        const float2 prev = B.kps[k];
        float px = prev.x * scale, py = prev.y * scale;
        if (level == G.top) {  // flags == 0, lk2d.cpp:51-56
          nx = px; ny = py;
          B.status[k] = 1;
          B.err[k] = 0.f;
        } else {
          float2 q = B.nxt[k];
          nx = q.x * 2.f; ny = q.y * 2.f;
        }
        B.nxt[k] = make_float2(nx, ny);  // lk2d.cpp:61
        px -= half; py -= half;
        const int ipx = cv_floor_d(px), ipy = cv_floor_d(py);
        if (ipx < -WIN || ipx >= lw || ipy < -WIN || ipy >= lh) {  // lk2d.cpp:91-99
          if (level == 0) { B.status[k] = 0; B.err[k] = 0.f; }
        } else {
          const Wts w = lk_weights(px - ipx, py - ipy);
          lk_sample_u8(I, st, ipx, ipy, w, Iw, sI, qI);
          float iA11 = 0, iA12 = 0, iA22 = 0;
          {  // derivative window + structure tensor, lk2d.cpp:117-149 (sums in (y, x) order)
            int tx[WIN], ty[WIN];
            const short2* row = dI + (ptrdiff_t)ipy * st + ipx;
            short2 nx2[WIN + 1];
#pragma unroll
            for (int x = 0; x <= WIN; x++) nx2[x] = row[x];
#pragma unroll 1
            for (int r = 0; r <= WIN; r++) {
              short2 cr[WIN + 1];
#pragma unroll
              for (int x = 0; x <= WIN; x++) cr[x] = nx2[x];
              row += st;
              if (r < WIN) {
#pragma unroll
                for (int x = 0; x <= WIN; x++) nx2[x] = row[x];
              }
              short2 pv = cr[0];
#pragma unroll
              for (int x = 0; x < WIN; x++) {
                short2 cv = cr[x + 1];
                int t0 = pv.x * w.w00 + cv.x * w.w01, b0 = pv.x * w.w10 + cv.x * w.w11;
                int t1 = pv.y * w.w00 + cv.y * w.w01, b1 = pv.y * w.w10 + cv.y * w.w11;
                if (r > 0) {
                  int ixv = LM_DESCALE(tx[x] + b0, W_BITS), iyv = LM_DESCALE(ty[x] + b1, W_BITS);
                  dIw[((r - 1) * WIN + x) * 32] = (unsigned)(ixv + 32768) | ((unsigned)(iyv + 32768) << 16);
                  // |ixv|, |iyv| <= 4080 (Scharr of u8, convex blend): the products are below 2^24, so the
                  // float product of the converted factors IS (float)(ixval * ixval) -- no I2F needed
                  const float fx = __int_as_float(0x4B400000 + ixv) - 12582912.0f;
                  const float fy = __int_as_float(0x4B400000 + iyv) - 12582912.0f;
                  iA11 += fx * fx;
                  iA12 += fx * fy;
                  iA22 += fy * fy;
                }
                tx[x] = t0; ty[x] = t1;
                pv = cv;
              }
            }
          }
          A11 = iA11 * FLT_SCALE; A12 = iA12 * FLT_SCALE; A22 = iA22 * FLT_SCALE;
          D = A11 * A22 - A12 * A12;
          const float min_eig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * WIN * WIN);
          if (min_eig < P.min_eig || D < FLT_EPSILON) {  // lk2d.cpp:294-298
            if (level == 0) B.status[k] = 0;
          } else {
            D = 1.f / D;
            outx = nx; outy = ny;  // value of nextPts[ptidx] (untouched if the loop leaves at once)
            nx -= half; ny -= half;
            pdx = 0.f; pdy = 0.f;
            j = 0;
            phase = P.max_count > 0 ? 1 : 3;
          }
        }
      }
      __syncwarp();
    }
    if (phase == 3) {  // maxCount == 0: the loop body never runs, j == maxCount
      if (level == 0) B.status[k] = 0;
      phase = 0;
    }
    if (__ballot_sync(0xffffffffu, phase != 0) == 0) {
      if (exhausted) break;
      continue;
    }
    if (phase != 0) {
      // ---- one trip: an LK iteration (lk2d.cpp:321-418) or the final error (:429-478) ----
      const float sx = phase == 1 ? nx : outx - half, sy = phase == 1 ? ny : outy - half;
      const int isx = cv_floor_d(sx), isy = cv_floor_d(sy);
      const bool outside = phase == 1 ? (isx < -half || isx >= lw || isy < -half || isy >= lh)
                                      : (isx < -WIN || isx >= lw || isy < -WIN || isy >= lh);
      bool finished = false, ok = true;
      if (outside) {
        if (phase == 2) { B.status[k] = 0; phase = 0; }
        else { finished = true; ok = false; }  // lk2d.cpp:326-330
      } else {
        const Wts w = lk_weights(sx - isx, sy - isy);
        int sJ;
        long long qJ;
        lk_sample_u8(J, st, isx, isy, w, Jw, sJ, qJ);
        float alpha = 1.0f, beta = 0.0f;
        if (P.illum) lk_norm_params(sI, qI, sJ, qJ, alpha, beta);
        float ib1 = 0, ib2 = 0, errval = 0;
#pragma unroll 13
        for (int t = 0; t < NWIN; t++) {  // lk2d.cpp:364-376 and :467-476, each sum in (y, x) order
          float fi, fj, fdx, fdy;
          u16x2_to_float(IJw[t * 32], -12582912.0f, fi, fj);
          u16x2_to_float(dIw[t * 32], -(12582912.0f + 32768.0f), fdx, fdy);
          float diff = alpha * fj + beta - fi;
          ib1 += diff * fdx;
          ib2 += diff * fdy;
          errval += fabsf(diff);
        }
        if (phase == 2) {
          B.err[k] = errval * 1.f / (32 * WIN * WIN);
          B.status[k] = 1;
          phase = 0;
        } else {
          const float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
          const float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
          nx += dx; ny += dy;
          outx = nx + half; outy = ny + half;
          if ((double)dx * dx + (double)dy * dy <= P.eps2) {  // lk2d.cpp:405
            finished = true;
          } else if (j > 0 && fabsf(dx + pdx) < 0.01 && fabsf(dy + pdy) < 0.01) {  // lk2d.cpp:410-414
            outx -= dx * 0.5f; outy -= dy * 0.5f;
            finished = true;
          } else {
            pdx = dx; pdy = dy;
            if (++j == P.max_count) { finished = true; ok = false; }  // lk2d.cpp:422
          }
        }
      }
      if (finished) {
        B.nxt[k] = make_float2(outx, outy);
        if (level != 0) phase = 0;
        else if (ok) phase = 2;
        else { B.status[k] = 0; phase = 0; }
      }
    }
    __syncwarp();
  }
}

// ---- closest line, vote, topological filter -------------------------------------------------------
__device__ __forceinline__ float point_line_distance(float x, float y, const float4 e) {  // lm.cpp:21-41
  float v_x = e.z - e.x, v_y = e.w - e.y;
  float u_x = e.x - x, u_y = e.y - y;
  float t = -(v_x * u_x + v_y * u_y) / (v_x * v_x + v_y * v_y);
  if (t < 0) t = 0;
  else if (t > 1) t = 1;
  float d_x = t * v_x + u_x, d_y = t * v_y + u_y;
  return sqrtf(d_x * d_x + d_y * d_y);
}

__global__ void __launch_bounds__(256) lm_vote_kernel(const VplLine* __restrict__ lines, const int* __restrict__ counts,
                                                      int cap, LmBuffers B, LmParams P, int pstride) {
  extern __shared__ unsigned char smem[];
  float4* s_end = reinterpret_cast<float4*>(smem);          // cap: endpoints of the current lines
  int* s_r2c = reinterpret_cast<int*>(s_end + cap);          // cap
  int* s_viol = s_r2c + cap;                                 // cap
  __shared__ int s_match_num;
  const int p = blockIdx.x;
  const int fr = p * pstride, fc = fr + 1;
  const int n_ref = min(counts[fr], cap), n_cur = min(counts[fc], cap);
  const VplLine* Lr = lines + (size_t)fr * cap;
  const VplLine* Lc = lines + (size_t)fc * cap;
  int* r2c_out = B.r2c + (size_t)p * cap;
  if (n_ref == 0 || n_cur == 0) {  // Matching returns false, lm.cpp:621: nothing is matched
    for (int i = threadIdx.x; i < n_ref; i += blockDim.x) r2c_out[i] = -1;
    if (threadIdx.x == 0) B.matched[p] = 0;
    return;
  }
  for (int i = threadIdx.x; i < n_cur; i += blockDim.x)
    s_end[i] = make_float4(Lc[i].endpoint[0], Lc[i].endpoint[1], Lc[i].endpoint[2], Lc[i].endpoint[3]);
  if (threadIdx.x == 0) s_match_num = 0;
  __syncthreads();
  const int n_kp = B.n_kp[p];
  const size_t kb = (size_t)p * B.cap_kp;
  int* kp2line = B.kp2line + kb;
  // ClosestLine, lm.cpp:48-86
  for (int i = threadIdx.x; i < n_kp; i += blockDim.x) {
    int label = -1;
    if (B.status[kb + i] && !(B.err[kb + i] > P.klt_err)) {
      const float2 pt = B.nxt[kb + i];
      int min_idx = -1;
      float min_d = 1000000;
      for (int j = 0; j < n_cur; j++) {
        float d = point_line_distance(pt.x, pt.y, s_end[j]);
        if (d < min_d) { min_d = d; min_idx = j; }
      }
      if (min_d < P.closest) label = min_idx;
    }
    kp2line[i] = label;
  }
  __syncthreads();
  // Point2Line, lm.cpp:88-133: the most voted current line of each reference line (lowest index
  // among equals, as the reference's `count[j] > max_value` scan)
  const int* kp_start = B.kp_start + (size_t)p * (cap + 1);
  for (int i = threadIdx.x; i < n_ref; i += blockDim.x) {
    const int a = kp_start[i], kn = kp_start[i + 1] - a;
    int max_value = 0, max_idx = 0;
    for (int u = 0; u < kn; u++) {
      const int lu = kp2line[a + u];
      if (lu < 0) continue;
      int c = 0;
      for (int v = 0; v < kn; v++) c += kp2line[a + v] == lu;
      if (c > max_value || (c == max_value && lu < max_idx)) { max_value = c; max_idx = lu; }
    }
    int m = -1;
    if (!(max_value <= 2 || (float)max_value / kn < P.ratio || Lc[max_idx].length > Lr[i].length * P.dist_ratio ||
          Lc[max_idx].length < Lr[i].length / P.dist_ratio))
      m = max_idx;
    s_r2c[i] = m;
    s_viol[i] = 0;
    if (m >= 0) atomicAdd(&s_match_num, 1);
  }
  __syncthreads();
  if (P.topo) {  // TopologicalFilter, lm.cpp:267-410
    for (int r1 = threadIdx.x; r1 < n_ref; r1 += blockDim.x) {
      const int c1 = s_r2c[r1];
      if (c1 < 0) continue;
      const double a_1 = Lr[r1].equation[0], b_1 = Lr[r1].equation[1], c_1 = Lr[r1].equation[2];
      double a_2 = Lc[c1].equation[0], b_2 = Lc[c1].equation[1], c_2 = Lc[c1].equation[2];
      if ((fabs(a_1 - a_2) + fabs(b_1 - b_2)) > (fabs(a_1 + a_2) + fabs(b_1 + b_2))) { a_2 = -a_2; b_2 = -b_2; c_2 = -c_2; }
      const double n1 = sqrt(a_1 * a_1 + b_1 * b_1), n2 = sqrt(a_2 * a_2 + b_2 * b_2);
      for (int r2 = 0; r2 < n_ref; r2++) {
        if (r2 == r1) continue;
        const int c2 = s_r2c[r2];
        if (c2 < 0) continue;
        const float lr2 = Lr[r2].length;
        if (fabsf(lr2 - Lc[c2].length) / lr2 > P.topo_len) continue;
        // SidenessCheck, lm.cpp:412-446
        const double px_1 = Lr[r2].center[0], py_1 = Lr[r2].center[1];
        const double px_2 = Lc[c2].center[0], py_2 = Lc[c2].center[1];
        const float d1 = (float)((px_1 * a_1 + py_1 * b_1 + c_1) / n1);
        const float d2 = (float)((px_2 * a_2 + py_2 * b_2 + c_2) / n2);
        if (d1 * d2 < 0 && fabsf(d1) > P.topo_dist && fabsf(d2) > P.topo_dist) {
          atomicAdd(&s_viol[r1], 1);
          atomicAdd(&s_viol[r2], 1);
        }
      }
    }
    __syncthreads();
    float threshold = P.topo_viol * (s_match_num - 1);
    if (threshold < 2) threshold = 2;
    for (int i = threadIdx.x; i < n_ref; i += blockDim.x)
      if (s_viol[i] > threshold) s_r2c[i] = -1;  // lm.cpp:656-665
    __syncthreads();
  }
  int matched = 0;
  for (int i = threadIdx.x; i < n_ref; i += blockDim.x) {
    r2c_out[i] = s_r2c[i];
    matched += s_r2c[i] >= 0;
  }
  // block count of the matches (for the caller's statistics)
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if (matched) atomicAdd(&s_cnt, matched);
  __syncthreads();
  if (threadIdx.x == 0) B.matched[p] = s_cnt;
}

}  // namespace

void launch_klt_pyramid(const uint8_t* img, uint8_t* pyr, short2* deriv, const KltGeom& G, int w, int h, int batch,
                        cudaStream_t st) {
  {
    dim3 grid((G.stride[0] / 4 + 127) / 128, h + 2 * G.pad, batch);
    klt_level0_kernel<<<grid, 128, 0, st>>>(img, pyr, G, w, h);
  }
  for (int l = 1; l <= G.top; l++) {
    dim3 grid((G.stride[l] + 127) / 128, G.h[l] + 2 * G.pad, batch);
    klt_pyrdown_kernel<<<grid, 128, 0, st>>>(pyr, G, l);
  }
  for (int l = 0; l <= G.top; l++) {
    dim3 grid((G.stride[l] / 4 + 127) / 128, G.h[l] + 2 * G.pad, batch);
    klt_scharr_kernel<<<grid, 128, 0, st>>>(pyr, deriv, G, l);
  }
}

void launch_lm_anchors(const VplLine* lines, const int* counts, int cap, const LmBuffers& B, const LmParams& P,
                       int pstride, int n_pairs, cudaStream_t st) {
  lm_anchor_kernel<<<n_pairs, 256, 0, st>>>(lines, counts, cap, B, P, pstride);
}

void launch_klt_track(const uint8_t* pyr, const short2* deriv, const KltGeom& G, const LmBuffers& B, const LmParams& P,
                      int pstride, int n_pairs, cudaStream_t st) {
  // per device, so set on every batch (a process may drive several GPUs)
  cudaFuncSetAttribute(klt_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrackSmem);
  cudaFuncSetAttribute(klt_track_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaMemsetAsync(B.queue, 0, kKltMaxLevels * sizeof(int), st);
  lm_pair_offsets_kernel<<<1, 1024, 0, st>>>(B.n_kp, n_pairs, B.pair_off);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int warps = sms * 5;  // 5 resident CTAs of 43 KB per SM
  static int refill = 0;      // idle lanes that trigger a refill (VPL_KLT_REFILL: tuning experiments)
  if (!refill) {
    const char* e = getenv("VPL_KLT_REFILL");
    refill = e ? atoi(e) : 16;  // measured on B200: 4 -> 25.8k, 8 -> 29.2k, 16 -> 32.2k, 22 -> 31.9k, 28 -> 29.3k, 32 -> 24.6k frames/s (E2)
    if (refill < 1 || refill > 32) refill = 16;
  }
  for (int level = G.top; level >= 0; level--)
    klt_track_kernel<<<warps, 32, kTrackSmem, st>>>(pyr, deriv, G, B, P, level, pstride, n_pairs, refill);
}

void launch_lm_vote(const VplLine* lines, const int* counts, int cap, const LmBuffers& B, const LmParams& P, int pstride,
                    int n_pairs, cudaStream_t st) {
  size_t smem = (size_t)cap * (sizeof(float4) + 2 * sizeof(int));
  // static + dynamic shared memory above 48 KB needs the opt-in (cap 2048 is exactly 48 KB of dynamic)
  if (smem > 40 * 1024) cudaFuncSetAttribute(lm_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lm_vote_kernel<<<n_pairs, 256, smem, st>>>(lines, counts, cap, B, P, pstride);
}

}  // namespace vpl
