// edlines.cu -- the reference's real line detector on the device (sm_100a).
//
// Replaces EDLineDetector::EDline (/root/reference/line_matching/src/edline_detector.cpp:1176;
// cited below as ed.cpp:line, ed.h = edline_detector.h), which the tracker calls through
// edline_detect (feature_tracker/src/line_feature_tracker.cpp:87, :315-321).  CPU restatement:
// oracle/orc_edlines.c (pinned bit for bit against the reference's own code).
//
// Stages (batch of B frames, all of them one launch over the whole batch):
//   ed_grad_anchor_kernel  Sobel pair, |dx|+|dy| -> threshold -> /4 (round half even) + direction bit
//                     (one u16 per pixel) and the anchor test on the scan grid -> column-major
//                     bitmap, one tiled pass over the image                          (ed.cpp:125-164)
//   ed_walk_kernel    smart routing.  Chains are claimed in anchor order and a walk stops at any
//                     earlier edge pixel, so it is sequential per frame: ONE WARP PER FRAME.  The warp
//                     expands the anchor bitmap 32 words at a time, tests 32 anchors for "already
//                     an edge pixel" with one gather, walks in lockstep (3 independent u16 loads
//                     per step: gradient, direction and edge mark share the word) while its lanes
//                     prefetch a strip of rows ahead of the walker, and re-packs each kept chain
//                                                                                    (ed.cpp:191-706)
//   ed_fit_kernel     ONE WARP PER EDGE CHAIN: least-squares fit of the first minLineLen pixels
//                     (integer sums, exact), extension 32 pixels per step with a ballot for the
//                     "4 consecutive outliers" rule, refits, Helmholtz validation   (ed.cpp:983-1173)
//   ed_compact_kernel lines into (chain, position) order -- the order the reference gives on one
//                     thread (its multi-threaded order is a race, ed.cpp:1081-1083)
// Everything is integer or IEEE double/float with -fmad=false, so the output equals the oracle's
// byte for byte (atan2/exp/log10/pow only feed threshold tests that sit far from their thresholds).
#include <float.h>

#include "vpl_common.cuh"
#include "vpl_tma.cuh"

namespace vpl {

namespace {

constexpr unsigned kG = 0x01ffu;     // (|dx|+|dy|)/4 <= 510
constexpr unsigned kEdge = 0x2000u;  // edge mark (pEdgeImg)
constexpr unsigned kDir = 0x8000u;   // set = Horizontal (|dx| < |dy|), ed.cpp:5-6, :136
enum { UP = 1, RIGHT = 2, DOWN = 3, LEFT = 4 };  // ed.cpp:7-10
constexpr int kTryTime = 6;                      // ed.cpp:11
constexpr int kSkipEdgePoint = 2;                // ed.cpp:12
constexpr unsigned kSearchMaxLen = 160;          // longest minLineLen handled by the lane-parallel window search

__device__ __forceinline__ unsigned gmap_value(short2 d, int grad_thresh) {
  int ax = abs((int)d.x), ay = abs((int)d.y);
  int s = ax + ay;
  if (!(s > grad_thresh + 1)) s = 0;  // threshold(THRESH_TOZERO, gradienThreshold_ + 1), ed.cpp:133
  int q = s >> 2, r = s & 3;          // Mat / 4 = convertTo(alpha .25): cvRound, ties to even
  if (r == 3 || (r == 2 && (q & 1))) q++;
  return (unsigned)q | (ax < ay ? kDir : 0u);
}

// Sobel pair, gradient map and anchor bitmap in one pass over the (already smoothed) image
// (ed.cpp:125-164).  A CTA takes a 64x16 tile: the image tile with a 2-pixel REFLECT_101 halo goes to
// shared memory, the map is computed on the tile plus a 1-pixel ring (an anchor compares with its
// neighbours' map values), and only grid points inside the tile are tested.
// Bit (ix*nH + iy) of the frame's bitmap = anchor at (1 + ix*scan, 1 + iy*scan): the bitmap is in
// the reference's column-major visiting order.
constexpr int GT_W = 64, GT_H = 16, GT_THREADS = 256;
constexpr int GT_SW = GT_W + 8;  // shared image row: columns x0-4 .. x0+GT_W+3 (word aligned)

template <int PITCH>
__device__ __forceinline__ short2 sobel_at(const uint8_t* p) {  // p -> centre pixel in the image tile of row pitch PITCH
  const int a = p[-PITCH - 1], b = p[-PITCH], c = p[-PITCH + 1], d = p[-1], e = p[1];
  const int g = p[PITCH - 1], h = p[PITCH], i = p[PITCH + 1];
  return make_short2((short)((c + 2 * e + i) - (a + 2 * d + g)), (short)((g + 2 * h + i) - (a + 2 * b + c)));
}

// TMA = true: the image tile arrives through the TMA unit (box x0-16 .. x0+79, 96 bytes wide, zero-filled outside the
// image, then reflected); false: per-thread loads (images whose pitch is not a multiple of 16 bytes).
constexpr int GT_TP = 96, GT_TX = 16;
template <bool TMA>
__global__ void __launch_bounds__(GT_THREADS) ed_grad_anchor_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                    const uint8_t* __restrict__ img,
                                                                    short2* __restrict__ grad,
                                                                    uint16_t* __restrict__ gmap,
                                                                    unsigned* __restrict__ bitmap,
                                                                    int* __restrict__ n_anchor, EdGeom G,
                                                                    int grad_thresh, int anchor_thresh) {
  constexpr int P = TMA ? GT_TP : GT_SW;   // row pitch of the image tile
  constexpr int X = TMA ? GT_TX : 4;       // tile column of image column x0
  __shared__ __align__(128) uint8_t s_img[GT_H + 4][P];  // rows y0-2 .. y0+GT_H+1
  __shared__ __align__(8) uint16_t s_g[GT_H + 2][GT_W + 4];  // map at (x0-1 .. x0+GT_W, y0-1 .. y0+GT_H), column c = x - x0 + 1
  __shared__ __align__(8) unsigned long long s_bar;
  const int f = blockIdx.z, w = G.w, h = G.h;
  const int x0 = blockIdx.x * GT_W, y0 = blockIdx.y * GT_H;
  const bool vec_ok = (w & 3) == 0;
  if (TMA) {
    tma_load_box_3d(&tmap, &s_img[0][0], &s_bar, x0 - GT_TX, y0 - 2, f, (GT_H + 4) * GT_TP);
    if (x0 - 4 < 0 || x0 + GT_W + 4 > w || y0 - 2 < 0 || y0 + GT_H + 2 > h)
      tma_reflect_fix(&s_img[0][0], GT_TP, GT_H + 4, GT_TX - 4, GT_SW, x0 - GT_TX, y0 - 2, w, h, GT_THREADS);
  } else {
    const uint8_t* src = img + (size_t)f * w * h;
    // image tile: 18 words per row x 20 rows; whole words inside the image are loaded as such
    for (int i = threadIdx.x; i < (GT_H + 4) * (GT_SW / 4); i += GT_THREADS) {
      const int r = i / (GT_SW / 4), cw = i - r * (GT_SW / 4);
      const int y = refl101(y0 + r - 2, h), x = x0 - 4 + 4 * cw;
      unsigned v;
      if (vec_ok && x >= 0 && x + 3 < w) {
        v = *reinterpret_cast<const unsigned*>(src + (size_t)y * w + x);
      } else {
        const uint8_t* row = src + (size_t)y * w;
        v = (unsigned)row[refl101(x, w)] | ((unsigned)row[refl101(x + 1, w)] << 8) | ((unsigned)row[refl101(x + 2, w)] << 16) |
            ((unsigned)row[refl101(x + 3, w)] << 24);
      }
      *reinterpret_cast<unsigned*>(&s_img[r][4 * cw]) = v;
    }
    __syncthreads();
  }
  const size_t fo = (size_t)f * w * h;
  {  // interior: 4 adjacent pixels per thread, 128-bit / 64-bit stores
    const int r = threadIdx.x >> 4, c4 = (threadIdx.x & 15) * 4;  // pixel (x0 + c4 .. +3, y0 + r)
    const int x = x0 + c4, y = y0 + r;
    short2 dv[4];
    unsigned gv[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      dv[k] = sobel_at<P>(&s_img[r + 2][c4 + k + X]);
      gv[k] = gmap_value(dv[k], grad_thresh);
      s_g[r + 1][c4 + k + 1] = (uint16_t)gv[k];
    }
    if (y < h) {
      if (vec_ok && x + 3 < w) {
        int4 o;
        o.x = *reinterpret_cast<int*>(&dv[0]); o.y = *reinterpret_cast<int*>(&dv[1]);
        o.z = *reinterpret_cast<int*>(&dv[2]); o.w = *reinterpret_cast<int*>(&dv[3]);
        *reinterpret_cast<int4*>(grad + fo + (size_t)y * w + x) = o;
        uint2 m;
        m.x = gv[0] | (gv[1] << 16); m.y = gv[2] | (gv[3] << 16);
        *reinterpret_cast<uint2*>(gmap + fo + (size_t)y * w + x) = m;
      } else {
        for (int k = 0; k < 4; k++)
          if (x + k < w) {
            grad[fo + (size_t)y * w + x + k] = dv[k];
            gmap[fo + (size_t)y * w + x + k] = (uint16_t)gv[k];
          }
      }
    }
  }
  // ring of map values around the tile (needed by the anchors on the tile's edge)
  if (threadIdx.x < 2 * (GT_W + 2) + 2 * GT_H) {
    int r, c;  // s_g coordinates
    const int t = threadIdx.x;
    if (t < GT_W + 2) { r = 0; c = t; }
    else if (t < 2 * (GT_W + 2)) { r = GT_H + 1; c = t - (GT_W + 2); }
    else if (t < 2 * (GT_W + 2) + GT_H) { r = 1 + t - 2 * (GT_W + 2); c = 0; }
    else { r = 1 + t - 2 * (GT_W + 2) - GT_H; c = GT_W + 1; }
    s_g[r][c] = (uint16_t)gmap_value(sobel_at<P>(&s_img[r + 1][c + X - 1]), grad_thresh);
  }
  __syncthreads();
  // anchors: grid points (1 + ix*scan, 1 + iy*scan) with 1 <= x <= w-2, 1 <= y <= h-2 inside this tile
  const int scan = G.scan;
  const int ix_lo = x0 <= 1 ? 0 : (x0 - 1 + scan - 1) / scan, iy_lo = y0 <= 1 ? 0 : (y0 - 1 + scan - 1) / scan;
  const int ix_hi = min(G.nW, (min(x0 + GT_W, w - 1) - 1 + scan - 1) / scan);  // exclusive
  const int iy_hi = min(G.nH, (min(y0 + GT_H, h - 1) - 1 + scan - 1) / scan);
  const int nx = ix_hi - ix_lo, ny = iy_hi - iy_lo;
  int found = 0;
  for (int i = threadIdx.x; i < nx * ny; i += GT_THREADS) {
    const int jy = i / nx, jx = i - jy * nx;
    const int ix = ix_lo + jx, iy = iy_lo + jy;
    const int c = 1 + ix * scan - x0 + 1, r = 1 + iy * scan - y0 + 1;  // position in s_g
    const unsigned v = s_g[r][c];
    const int gv = v & kG;
    int a, b;
    if (v & kDir) { a = s_g[r - 1][c] & kG; b = s_g[r + 1][c] & kG; }
    else { a = s_g[r][c - 1] & kG; b = s_g[r][c + 1] & kG; }
    if (gv >= a + anchor_thresh && gv >= b + anchor_thresh) {
      const int bit = ix * G.nH + iy;
      atomicOr(&bitmap[(size_t)f * G.bm_words + (bit >> 5)], 1u << (bit & 31));
      found++;
    }
  }
  if (found) atomicAdd(&n_anchor[f], found);
}

// ---- smart routing ---------------------------------------------------------------------------
struct WalkMem {
  int pfx, pfy;  // centre of the last prefetched strip
};

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// A walk is a chain of dependent loads; when it runs across rows every step touches a new cache
// line (DRAM latency per pixel).  The 32 lanes therefore keep a strip of 32 rows x ~100 columns
// ahead of the walker on its way into L1: one prefetch per lane, three columns, re-issued when the
// walker leaves the inner box of the strip.
__device__ __forceinline__ void ed_prefetch_strip(const uint16_t* __restrict__ g, int W, int H, int x, int y, int mx,
                                                  int my, int lane, WalkMem& wm) {
  if (abs(x - wm.pfx) > 24 || abs(y - wm.pfy) > 8) {
    wm.pfx = x + 16 * mx;
    wm.pfy = y + 8 * my;
    int ry = min(max(wm.pfy - 16 + lane, 0), H - 1);
    const uint16_t* row = g + (size_t)ry * W;
    prefetch_l1(row + min(max(wm.pfx - 40, 0), W - 1));
    prefetch_l1(row + min(max(wm.pfx, 0), W - 1));
    prefetch_l1(row + min(max(wm.pfx + 40, 0), W - 1));
  }
}

// One directional walk (ed.cpp:211-312 and its three copies), executed by the whole warp in
// lockstep on identical state (SIMT issues one instruction stream either way; the loads are
// broadcasts).  Every lane stores the edge mark it will read back later; lane 0 records the pixel.
// Returns the number of pixels recorded, or -1 if `cap` would be exceeded.
//
// The reference's four-way case analysis is folded into arithmetic on a direction code
// (0 right, 1 left, 2 down, 3 up; bit 1 = vertical, bit 0 = negative):
//   - a horizontal pixel reached by a vertical move (or the reverse) turns towards the side the last
//     move drifted to: the reference tests `x > lastX` / `y > lastY`, and lastX/lastY always hold the
//     previous pixel of the same walk when they are read, so that is the sign of the last step;
//   - the three candidates are idx + dmain + {dperp, 0, -dperp}: for a horizontal move dmain = +-1 and
//     candidate 1 is the upper one (dperp = -W), for a vertical move dmain = +-W and candidate 1 is
//     the right one (dperp = +1), exactly the reference's gValue1 / gValue2 / gValue3.
//
// (Tried and dropped: edge marks as a whole-frame bitmap in shared memory with a read-only gradient
// map -- 45 KB per walker cuts the walkers per SM from 28 to 4: 10 ms -> 39 ms per 4096 frames; and
// a 64 x 32 tile of the map in shared memory that follows the walker -- most walks are a few pixels
// long, so the 4 KB tile reloads cost more than the L2 round trips they save: 4.3 -> 5.5 ms per 512
// frames.  The loads after a mark store remain the top stall of this kernel.)
__device__ __noinline__ int ed_walk(uint16_t* __restrict__ g, int W, int H, unsigned x0, unsigned y0, int last_dir,
                                    uint32_t* __restrict__ out, int cap, WalkMem& wm, int lane) {
  int n = 0;
  int x = (int)x0, y = (int)y0;
  int idx = y * W + x;
  unsigned v = g[idx];
  int ld = last_dir == RIGHT ? 0 : last_dir == LEFT ? 1 : last_dir == DOWN ? 2 : 3;
  int ddx = 0, ddy = 0;  // last step (never read before the first step has set it)
  while ((v & kG) != 0 && !(v & kEdge)) {
    if (n >= cap) return -1;
    g[idx] = (uint16_t)(v | kEdge);
    if (lane == 0) out[n] = (unsigned)x | ((unsigned)y << 16);
    n++;
    const int h = (int)(v >> 15) & 1;  // 1 = horizontal pixel
    int go = ld;
    if (h == (ld >> 1)) go = h ? (ddx > 0 ? 0 : 1) : (ddy > 0 ? 2 : 3);  // ed.cpp:217-223 / :264-270
    const int sgn = 1 - 2 * (go & 1);
    const bool hz = go < 2;
    // image border: the three candidates must exist (ed.cpp:227, :245, :274, :292)
    const int a = hz ? x : y, amax = hz ? W : H, bq = hz ? y : x, bmax = hz ? H : W;
    if ((unsigned)(a + sgn) >= (unsigned)amax || (unsigned)(bq - 1) >= (unsigned)(bmax - 2)) break;
    const int dmain = hz ? sgn : sgn * W, dperp = hz ? -W : 1;
    const int i2 = idx + dmain;
    const unsigned v1 = g[i2 + dperp], v2 = g[i2], v3 = g[i2 - dperp];
    if ((n & 3) == 0) ed_prefetch_strip(g, W, H, x, y, hz ? sgn : 0, hz ? 0 : sgn, lane, wm);
    const unsigned g1 = v1 & 0xffu, g2 = v2 & 0xffu, g3 = v3 & 0xffu;  // the (unsigned char) casts, ed.cpp:231-233
    const int sel = (g1 >= g2 && g1 >= g3) ? 1 : ((g3 >= g2 && g3 >= g1) ? -1 : 0);
    idx = i2 + sel * dperp;
    v = sel > 0 ? v1 : (sel < 0 ? v3 : v2);
    ddx = hz ? sgn : sel;
    ddy = hz ? -sel : sgn;
    x += ddx; y += ddy;
    ld = go;
  }
  return n;
}

__global__ void __launch_bounds__(32) ed_walk_kernel(EdBuffers B, EdGeom G, int min_len, int batch) {
  __shared__ uint32_t s_anchor[1024];  // anchors of the current 32 bitmap words, in visiting order
  const int f = blockIdx.x;
  if (f >= batch) return;
  const int lane = threadIdx.x;
  uint16_t* g = B.gmap + (size_t)f * G.w * G.h;
  const unsigned* bm = B.bitmap + (size_t)f * G.bm_words;
  uint32_t* first = B.first + (size_t)f * G.part_cap;
  uint32_t* second = B.second + (size_t)f * G.part_cap;
  uint32_t* xy = B.xy + (size_t)f * 2 * G.cap_px;
  uint32_t* sid = B.sid + (size_t)f * (G.cap_edges + 2);
  int status = 1;
  if (B.n_anchor[f] > G.cap_px) status = -1;  // ed.cpp:166-169
  int off1 = 0, off2 = 0, n_edge = 0, k = 0;
  WalkMem wm = {-1000000, -1000000};
  for (int base = 0; status == 1 && base < G.bm_words; base += 32) {
    // expand the set bits of 32 words into the anchor list (column-major order = bit order)
    unsigned word = (base + lane < G.bm_words) ? bm[base + lane] : 0u;
    int cnt = __popc(word), incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    {
      int o = incl - cnt;
      unsigned wv = word;
      while (wv) {
        int b = __ffs(wv) - 1;
        wv &= wv - 1;
        int cand = ((base + lane) << 5) + b;
        int ix = cand / G.nH, iy = cand - ix * G.nH;
        s_anchor[o++] = (unsigned)(1 + ix * G.scan) | ((unsigned)(1 + iy * G.scan) << 16);
      }
    }
    __syncwarp();
    for (int a0 = 0; status == 1 && a0 < total; a0 += 32) {
      // 32 anchors at a time: one gather tells which are already edge pixels (ed.cpp:195); the
      // gather is repeated for the remaining ones after every walk, since a walk sets marks
      const bool have = a0 + lane < total;
      const unsigned axy = have ? s_anchor[a0 + lane] : 0u;
      const int aidx = (int)(axy >> 16) * G.w + (int)(axy & 0xffff);
      unsigned todo = __ballot_sync(0xffffffffu, have && !(g[aidx] & kEdge));
      while (todo && status == 1) {
        const int src = __ffs(todo) - 1;
        const unsigned a = __shfl_sync(0xffffffffu, axy, src);
        const unsigned x = a & 0xffff, y = a >> 16;
        const int idx = (int)y * G.w + (int)x;
        const bool horizontal = (g[idx] & kDir) != 0;
        const int len1 = ed_walk(g, G.w, G.h, x, y, horizontal ? RIGHT : DOWN, first, G.part_cap, wm, lane);
        g[idx] = (uint16_t)(g[idx] & ~kEdge);  // the anchor is walked again, ed.cpp:317 / :533
        const int len2 = ed_walk(g, G.w, G.h, x, y, horizontal ? LEFT : UP, second, G.part_cap, wm, lane);
        __syncwarp();
        // anchors after this one that the two walks have just covered drop out
        const unsigned later = src == 31 ? 0u : (0xffffffffu << (src + 1));
        todo = __ballot_sync(0xffffffffu, have && !(g[aidx] & kEdge)) & todo & later;
        if (len1 < 0 || len2 < 0) { status = -1; break; }
        if (len1 + len2 < min_len + 1) continue;  // short edge: records dropped, marks stay, ed.cpp:641-643
        off1 += len1; off2 += len2;
        if (off1 > G.cap_px || off2 > G.cap_px || n_edge + 1 > G.cap_edges) { status = -1; break; }  // ed.cpp:655-666
        // re-pack: first part reversed, then the second part without the anchor, ed.cpp:687-702
        if (lane == 0) sid[n_edge] = k;
        for (int i = lane; i < len1; i += 32) xy[k + i] = first[len1 - 1 - i];
        for (int i = lane; i < len2 - 1; i += 32) xy[k + len1 + i] = second[1 + i];
        k += len1 + (len2 > 0 ? len2 - 1 : 0);
        n_edge++;
        __syncwarp();
      }
    }
    __syncwarp();
  }
  if (!(off1 && off2)) status = -1;  // ed.cpp:667 "lines not found"
  if (lane == 0) {
    if (status == 1) sid[n_edge] = k;
    B.n_chain[f] = status == 1 ? n_edge : 0;
    B.n_px[f] = status == 1 ? k : 0;
    B.status[f] = status;
  }
}

// ---- per-chain line extraction ------------------------------------------------------------------
__device__ __forceinline__ bool ed_double_equal(double a, double b) {  // ed.h:171-187
  if (a == b) return true;
  double abs_diff = fabs(a - b), aa = fabs(a), bb = fabs(b);
  double abs_max = aa > bb ? aa : bb;
  if (abs_max < DBL_MIN) abs_max = DBL_MIN;
  return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}

// nfa(n, k, p, logNT), ed.h:275-348.  log_gamma of integer arguments comes from the table the
// host built with the same Lanczos/Windschitl formulas (ed.h:210-240, capi.cu host_log_gamma).
__device__ __noinline__ double ed_nfa(int n, int k, double p, double logNT, const double* __restrict__ lgam) {
  if (n == 0 || k == 0) return -logNT;
  if (n == k) return -logNT - (double)n * log10(p);
  double p_term = p / (1.0 - p);
  double log1term = lgam[n + 1] - lgam[k + 1] - lgam[n - k + 1] + (double)k * log(p) + (double)(n - k) * log(1.0 - p);
  double term = exp(log1term);
  if (ed_double_equal(term, 0.0)) {
    if ((double)k > (double)n * p) return -log1term / 2.30258509299404568402 - logNT;
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
      if (err < 0.1 * fabs(-log10(bin_tail) - logNT) * bin_tail) break;
    }
  }
  return -log10(bin_tail) - logNT;
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums of the normal equations over chain pixels [s, e): exact integers, cast to float once
// (= cv::gemm's double accumulators on integer-valued floats, ed.cpp:757-758 / :843-844)
__device__ __forceinline__ void ed_normal_sums(const uint32_t* __restrict__ xy, unsigned s, unsigned e, bool horizontal,
                                               int lane, float ata[4], float atv[2]) {
  long long suu = 0, suv = 0;
  int su = 0, sv = 0;
  for (unsigned i = s + lane; i < e; i += 32) {
    unsigned p = xy[i];
    int x = p & 0xffff, y = p >> 16;
    int u = horizontal ? x : y, v = horizontal ? y : x;
    suu += (long long)u * u; suv += (long long)u * v; su += u; sv += v;
  }
  suu = warp_sum_ll(suu); suv = warp_sum_ll(suv); su = warp_sum_i(su); sv = warp_sum_i(sv);
  ata[0] = (float)(double)suu; ata[1] = (float)(double)su; ata[2] = ata[1]; ata[3] = (float)(double)(int)(e - s);
  atv[0] = (float)(double)suv; atv[1] = (float)(double)sv;
}
__device__ __forceinline__ void ed_fit_solve(const float a[4], const float v[2], double& eq0, double& eq1) {  // ed.cpp:761-764
  double coef = 1.0 / ((double)a[0] * (double)a[3] - (double)a[1] * (double)a[2]);
  eq0 = coef * ((double)a[3] * (double)v[0] - (double)a[1] * (double)v[1]);
  eq1 = coef * ((double)a[0] * (double)v[1] - (double)a[2] * (double)v[0]);
}

__global__ void __launch_bounds__(128) ed_fit_kernel(EdBuffers B, EdGeom G, int min_len, double thr,
                                                     const short2* __restrict__ grad, const double* __restrict__ lgam) {
  __shared__ double s_sq[4][32];
  __shared__ uint32_t s_pts[4][64 + kSearchMaxLen];
  const int f = blockIdx.y;
  if (B.status[f] != 1) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_chain = B.n_chain[f];
  const uint16_t* g = B.gmap + (size_t)f * G.w * G.h;
  const short2* gr = grad + (size_t)f * G.w * G.h;
  const uint32_t* xy = B.xy + (size_t)f * 2 * G.cap_px;
  const uint32_t* sid = B.sid + (size_t)f * (G.cap_edges + 2);
  VplLine* slots = B.slots + (size_t)f * G.nslots;
  uint8_t* valid = B.slot_valid + (size_t)f * G.nslots;
  const unsigned m = (unsigned)min_len;
  double* sq = s_sq[warp];
  uint32_t* spts = s_pts[warp];

  for (int e = blockIdx.x * 4 + warp; e < n_chain; e += gridDim.x * 4) {
    unsigned S = sid[e];
    const unsigned E = sid[e + 1];
    double eq0 = 0, eq1 = 0, fit_err = 0;
    float ata[4], atv[2];
    while (E > S + m) {  // ed.cpp:987
      bool horizontal = false;
      if (m <= kSearchMaxLen) {
        // find an initial segment, ed.cpp:989-995.  The reference tries the windows starting at
        // S, S+2, S+4, ... one after the other and stops at the first whose fit error passes; the
        // windows are independent, so the 32 lanes try 32 consecutive starts at once (each lane its
        // own exact integer sums and its own in-order residual sum) and the lowest passing lane wins.
        bool found = false;
        while (E > S + m) {
          const unsigned span = min(E - S, 62u + m);
          for (unsigned i = lane; i < span; i += 32) spts[i] = xy[S + i];
          __syncwarp();
          const unsigned o = 2u * lane;
          const bool cand = E > S + o + m;
          bool hz = false;
          double q0 = 0, q1 = 0, ferr = 0;
          float a[4] = {0, 0, 0, 0}, v2[2] = {0, 0};
          if (cand) {
            const unsigned p0 = spts[o];
            hz = (g[(p0 >> 16) * G.w + (p0 & 0xffff)] & kDir) != 0;
            int su = 0, sv = 0;
            long long luu = 0, luv = 0;
            for (unsigned i = 0; i < m; i++) {
              const unsigned p = spts[o + i];
              const int x = p & 0xffff, y = p >> 16;
              const int u = hz ? x : y, vv = hz ? y : x;
              luu += (long long)u * u; luv += (long long)u * vv; su += u; sv += vv;
            }
            a[0] = (float)(double)luu; a[1] = (float)(double)su; a[2] = a[1]; a[3] = (float)(double)(int)m;
            v2[0] = (float)(double)luv; v2[1] = (float)(double)sv;
            ed_fit_solve(a, v2, q0, q1);
            double err = 0;  // squared residuals in pixel order, ed.cpp:767-770
            for (unsigned i = 0; i < m; i++) {
              const unsigned p = spts[o + i];
              const double x = (double)(p & 0xffff), y = (double)(p >> 16);
              const double c = hz ? y - x * q0 - q1 : x - y * q0 - q1;
              err += c * c;
            }
            ferr = sqrt(err);
          }
          const unsigned okm = __ballot_sync(0xffffffffu, cand && ferr <= thr);
          const unsigned candm = __ballot_sync(0xffffffffu, cand);
          __syncwarp();
          if (okm) {
            const int l = __ffs(okm) - 1;
            S += 2u * (unsigned)l;
            eq0 = __shfl_sync(0xffffffffu, q0, l); eq1 = __shfl_sync(0xffffffffu, q1, l);
            fit_err = __shfl_sync(0xffffffffu, ferr, l);
#pragma unroll
            for (int t = 0; t < 4; t++) ata[t] = __shfl_sync(0xffffffffu, a[t], l);
            atv[0] = __shfl_sync(0xffffffffu, v2[0], l); atv[1] = __shfl_sync(0xffffffffu, v2[1], l);
            horizontal = __shfl_sync(0xffffffffu, (int)hz, l) != 0;
            found = true;
            break;
          }
          S += 2u * (unsigned)__popc(candm);  // every candidate of this round failed
        }
        if (!found) break;  // the last attempt failed: lineFitErr > threshold, ed.cpp:996
      } else {
      while (E > S + m) {  // find an initial segment, ed.cpp:989-995 (window too long for the lane-parallel search)
        unsigned p0 = xy[S];
        horizontal = (g[(p0 >> 16) * G.w + (p0 & 0xffff)] & kDir) != 0;
        ed_normal_sums(xy, S, S + m, horizontal, lane, ata, atv);
        ed_fit_solve(ata, atv, eq0, eq1);
        double err = 0;  // sum of squared residuals in pixel order, ed.cpp:767-770
        for (unsigned base = 0; base < m; base += 32) {
          unsigned i = base + lane;
          double c2 = 0;
          if (i < m) {
            unsigned p = xy[S + i];
            double x = (double)(p & 0xffff), y = (double)(p >> 16);
            double c = horizontal ? y - x * eq0 - eq1 : x - y * eq0 - eq1;
            c2 = c * c;
          }
          sq[lane] = c2;
          __syncwarp();
          int cnt = min(32u, m - base);
          for (int j = 0; j < cnt; j++) err += sq[j];
          __syncwarp();
          // the partial sums only grow: once sqrt(partial) > thr the verdict is final
          if (sqrt(err) > thr) break;
        }
        fit_err = sqrt(err);
        if (fit_err <= thr) break;
        S += kSkipEdgePoint;
      }
      }
      if (fit_err > thr) break;  // ed.cpp:996
      // extend, ed.cpp:1008-1039 / :1090-1120
      const unsigned S0 = S;
      double coef1 = 0;
      bool extended = true, first_try = true;
      int tries = 0, outliers = 0;
      unsigned new_s = 0;
      while (extended) {
        tries++;
        if (first_try) {
          first_try = false;
          S += m;
        } else {  // incremental refit with the pixels [new_s, S), ed.cpp:805-891
          float a2[4], v2[2];
          ed_normal_sums(xy, new_s, S, horizontal, lane, a2, v2);
#pragma unroll
          for (int i = 0; i < 4; i++) ata[i] = ata[i] + a2[i];
          atv[0] = atv[0] + v2[0]; atv[1] = atv[1] + v2[1];
          ed_fit_solve(ata, atv, eq0, eq1);
        }
        coef1 = 1 / sqrt(eq0 * eq0 + 1);
        outliers = 0;
        new_s = S;
        bool stop = false;
        while (E > S && !stop) {
          unsigned i = S + lane;
          bool in = i < E, far = false;
          if (in) {
            unsigned p = xy[i];
            unsigned X = p & 0xffff, Y = p >> 16;
            double d = horizontal ? fabs(eq0 * X - Y + eq1) * coef1 : fabs(X - eq0 * Y - eq1) * coef1;
            far = d > thr;
          }
          unsigned mo = __ballot_sync(0xffffffffu, far);
          int nin = __popc(__ballot_sync(0xffffffffu, in));
          // bits 0..2: the run of outliers carried in (right-aligned against bit 3 = pixel 0)
          unsigned long long wd = ((unsigned long long)mo << 3) | (unsigned long long)(((1u << outliers) - 1u) << (3 - outliers));
          unsigned long long r4 = wd & (wd >> 1) & (wd >> 2) & (wd >> 3);
          if (r4) {  // fourth consecutive outlier at pixel j: numOfOutlier > 3, ed.cpp:1027
            int j = __ffsll((long long)r4) - 1;
            S += (unsigned)j + 1;
            outliers = 4;
            stop = true;
          } else {
            S += (unsigned)nin;
            int top = nin + 2;  // highest bit in use
            outliers = __clzll((long long)~(wd << (63 - top)));
          }
        }
        S -= (unsigned)outliers;  // pop the trailing outliers, ed.cpp:1034
        if (!(S != new_s && tries < kTryTime)) extended = false;
      }
      double le0, le1, le2;
      if (horizontal) { le0 = eq0 * coef1; le1 = -1 * coef1; le2 = eq1 * coef1; }  // ed.cpp:1041-1044
      else { le0 = 1 * coef1; le1 = -eq0 * coef1; le2 = -eq1 * coef1; }            // ed.cpp:1122-1125
      // LineValidation, ed.cpp:893-958
      const int n = (int)(S - S0);
      int mgx = 0, mgy = 0;
      for (unsigned i = S0 + lane; i < S; i += 32) {
        unsigned p = xy[i];
        short2 d = gr[(p >> 16) * G.w + (p & 0xffff)];
        mgx += d.x; mgy += d.y;
      }
      mgx = warp_sum_i(mgx); mgy = warp_sum_i(mgy);
      bool ok = !(mgx == 0 && mgy == 0);
      float direction = 0.f;
      if (ok) {
        double adx = fabs(le1), ady = fabs(le0);
        if (mgx > 0 && mgy >= 0) direction = (float)atan2(-ady, adx);
        if (mgx <= 0 && mgy > 0) direction = (float)atan2(ady, adx);
        if (mgx < 0 && mgy <= 0) direction = (float)atan2(ady, -adx);
        if (mgx >= 0 && mgy < 0) direction = (float)atan2(-ady, -adx);
        double fd = fabs((double)direction);
        if (fd < 0.15 || VPL_PI - fd < 0.15)
          if (fabs(le2) < 10 || fabs((double)(unsigned)G.h - fabs(le2)) < 10) ok = false;
        if (ok && fabs(fd - VPL_PI * 0.5) < 0.15)
          if (fabs(le2) < 10 || fabs((double)(unsigned)G.w - fabs(le2)) < 10) ok = false;
      }
      if (ok) {
        int k = 0;
        for (unsigned i = S0 + lane; i < S; i += 32) {
          unsigned p = xy[i];
          short2 d = gr[(p >> 16) * G.w + (p & 0xffff)];
          double pd = atan2(-(double)d.x, (double)d.y);
          double dis = fabs((double)direction - pd);
          if (fabs(2 * VPL_PI - dis) < 0.392699 || dis < 0.392699) k++;
        }
        k = warp_sum_i(k);
        ok = ed_nfa(n, k, 0.125, G.logNT, lgam) > 0;
      }
      if (ok && lane == 0) {  // endpoints = projections of the first and last pixel, ed.cpp:1056-1079
        double a1 = le1 * le1, a2 = le0 * le0, a3 = le0 * le1, a4 = le2 * le0, a5 = le2 * le1;
        unsigned p = xy[S0];
        unsigned Px = p & 0xffff, Py = p >> 16;
        float x1 = (float)(a1 * Px - a3 * Py - a4), y1 = (float)(a2 * Py - a3 * Px - a5);
        p = xy[S - 1];
        Px = p & 0xffff; Py = p >> 16;
        float x2 = (float)(a1 * Px - a3 * Py - a4), y2 = (float)(a2 * Py - a3 * Px - a5);
        VplLine L;
        L.endpoint[0] = x1; L.endpoint[1] = y1; L.endpoint[2] = x2; L.endpoint[3] = y2;
        L.equation[0] = le0; L.equation[1] = le1; L.equation[2] = le2;
        L.center[0] = (float)((double)(x1 + x2) / 2.0);
        L.center[1] = (float)((double)(y1 + y2) / 2.0);
        double ddx = (double)(x2 - x1), ddy = (double)(y2 - y1);
        L.length = (float)sqrt(ddx * ddx + ddy * ddy);
        L.reserved = 0;
        // lines of a frame start >= minLineLen pixels apart, so S0 / minLineLen is a free slot
        // and slots are in (chain, position) order
        unsigned slot = S0 / m;
        slots[slot] = L;
        valid[slot] = 1;
      }
    }
  }
}

// one warp per frame: valid slots -> dense rows, order kept
__global__ void __launch_bounds__(32) ed_compact_kernel(EdBuffers B, EdGeom G, VplLine* __restrict__ out,
                                                        int* __restrict__ counts, int cap, int* __restrict__ overflow,
                                                        int batch) {
  const int f = blockIdx.x;
  if (f >= batch) return;
  const int lane = threadIdx.x;
  const VplLine* slots = B.slots + (size_t)f * G.nslots;
  const uint8_t* valid = B.slot_valid + (size_t)f * G.nslots;
  int n = 0;
  if (B.status[f] == 1) {
    int used = B.n_px[f] / G.min_len + 1;
    if (used > G.nslots) used = G.nslots;
    for (int base = 0; base < used; base += 32) {
      int i = base + lane;
      bool v = i < used && valid[i];
      unsigned mask = __ballot_sync(0xffffffffu, v);
      int pos = n + __popc(mask & ((1u << lane) - 1u));
      if (v && pos < cap) out[(size_t)f * cap + pos] = slots[i];
      n += __popc(mask);
    }
  }
  if (lane == 0) {
    if (n > cap) { *overflow = 1; }
    counts[f] = n > cap ? cap : n;
  }
}

}  // namespace

void launch_ed_grad_anchor(const uint8_t* img, short2* grad, const EdBuffers& B, const EdGeom& G, int grad_thresh,
                           int anchor_thresh, int batch, cudaStream_t st) {
  cudaMemsetAsync(B.bitmap, 0, (size_t)batch * G.bm_words * sizeof(unsigned), st);
  cudaMemsetAsync(B.n_anchor, 0, (size_t)batch * sizeof(int), st);
  dim3 grid((G.w + GT_W - 1) / GT_W, (G.h + GT_H - 1) / GT_H, batch);
  CUtensorMap tm;
  if (make_u8_frames_tmap(&tm, img, G.w, G.h, batch, GT_TP, GT_H + 4))
    ed_grad_anchor_kernel<true><<<grid, GT_THREADS, 0, st>>>(tm, img, grad, B.gmap, B.bitmap, B.n_anchor, G, grad_thresh, anchor_thresh);
  else {
    memset(&tm, 0, sizeof(tm));
    ed_grad_anchor_kernel<false><<<grid, GT_THREADS, 0, st>>>(tm, img, grad, B.gmap, B.bitmap, B.n_anchor, G, grad_thresh, anchor_thresh);
  }
}

void launch_ed_walk(const EdBuffers& B, const EdGeom& G, int batch, cudaStream_t st) {
  ed_walk_kernel<<<batch, 32, 0, st>>>(B, G, G.min_len, batch);

}

void launch_ed_fit(const EdBuffers& B, const EdGeom& G, double fit_thr, const short2* grad, const double* lgam,
                   int batch, cudaStream_t st) {
  cudaMemsetAsync(B.slot_valid, 0, (size_t)batch * G.nslots, st);
  dim3 grid(16, batch);  // 64 warps per frame over its chains
  ed_fit_kernel<<<grid, 128, 0, st>>>(B, G, G.min_len, fit_thr, grad, lgam);
}

void launch_ed_compact(const EdBuffers& B, const EdGeom& G, VplLine* out, int* counts, int cap, int* overflow,
                       int batch, cudaStream_t st) {
  ed_compact_kernel<<<batch, 32, 0, st>>>(B, G, out, counts, cap, overflow, batch);
}

}  // namespace vpl
