// lsd_engine.cu -- LSD region engine (sm_100a): region growing, region -> rectangle,
// refine (density test, re-grow with tighter tolerance, radius reduction).
//
// Restates the main loop of cv::LineSegmentDetectorImpl::flsd and region_grow /
// region2rect / get_theta / refine / reduce_region_radius (opencv imgproc lsd.cpp;
// SURVEY.md Appendix A.4-A.6; CPU restatement oracle/orc_lsd.c).  rect_improve /
// rect_nfa do not feed back into the `used` map, so they run afterwards in a
// separate, fully parallel kernel (lsd_nfa.cu) on the candidates emitted here.
//
// Execution model: ONE WARP PER (frame, octave).  The seed loop of a frame is
// sequential by definition (first-come pixel ownership, running float32 mean
// angle, double sums in list order), so the warp keeps exactly that order and
// uses its 32 lanes for everything that is order-free:
//   * seed scan: 32 ordered seeds per step, one gather each;
//   * growth: the 8 neighbours of up to four FIFO pixels are gathered with one 128-bit load per lane
//     (angle with the used flag in its sign bit, cos, sin, gradient differences);
//     acceptance is then resolved in lane order, which is exactly the sequential (FIFO, yy, xx) order;
//   * rectangle sums: per-entry products in parallel, the additions themselves in
//     list order (shuffle broadcast) so the double results equal the sequential
//     sums bit for bit;
//   * min/max projections, used-bit clearing, radius compaction: lane-parallel.
// Throughput comes from frame-level parallelism (thousands of warps in flight),
// not from speculation; results are bit-identical to the sequential algorithm.
//
// Compile with -fmad=false: the double/float expressions below mirror the CPU
// sequence operation by operation and must not be contracted.
#include "vpl_common.cuh"
#include "vpl_sincos.cuh"

#include <cstdlib>

namespace vpl {

#ifndef VPL_ENGINE_RING
#define VPL_ENGINE_RING 256
#endif
constexpr int RING = VPL_ENGINE_RING;  // per-warp FIFO window kept in shared memory

struct Eng {
  Pix* pix;            // {angle bits | used bit31, cosf, sinf, packed gradient differences} per pixel
  uint32_t* reg;       // region list: x | y << 16 per entry
  int* ring;  // shared memory, RING ints
  double* bc;    // shared memory, 32 doubles  } broadcast buffers for the in-order sums
  double2* bc2;  // shared memory, 32 double2  }
  int ws, hs;
  int lane;
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__device__ __forceinline__ double dist_d(double x1, double y1, double x2, double y2) {
  return sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
}
__device__ __forceinline__ double dist_sq_d(double x1, double y1, double x2, double y2) {
  return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1);
}
__device__ __forceinline__ double angle_diff_signed_d(double a, double b) {
  double d = a - b;
  while (d <= -VPL_PI) d += VPL_2PI;
  while (d > VPL_PI) d -= VPL_2PI;
  return d;
}

// isAligned on an angle already converted to radians
__device__ __forceinline__ bool aligned_rad(double a, double theta, double prec) {
  double n_theta = theta - a;
  if (n_theta < 0) n_theta = -n_theta;
  if (n_theta > VPL_3_2_PI) {
    n_theta -= VPL_2PI;
    if (n_theta < 0) n_theta = -n_theta;
  }
  return n_theta <= prec;
}

// ---------------------------------------------------------------------------
// region_grow (A.4).  Returns the region size; reg[0..n) holds the region in
// acceptance order; *reg_angle_out is the final running angle.
// ---------------------------------------------------------------------------
__device__ __noinline__ int region_grow(const Eng& e, int seed_xy, float seed_deg, double prec,
                                        double* reg_angle_out) {
  const int lane = e.lane, ws = e.ws, hs = e.hs;
  Pix* pix = e.pix;
  // seed: defined and currently unused
  if (lane == 0) {
    pix[(seed_xy >> 16) * ws + (seed_xy & 0xffff)].ang = __float_as_uint(seed_deg) | kUsedBit;
    e.reg[0] = (uint32_t)seed_xy;
    e.ring[0] = seed_xy;
  }
  // The running region angle is always (double)reg_deg * DEG2RAD with reg_deg the float32
  // fastAtan2 result, so the alignment test is first decided in float32 on the degrees and
  // only falls back to the exact double sequence inside a guard band around the thresholds
  // (float error on the difference is < 1e-4 deg; the band is 1e-2 deg): same decisions,
  // a fraction of the FP64 work.
  float reg_deg = seed_deg;
  float sumdx, sumdy;
  {
    double reg_angle0 = (double)seed_deg * VPL_DEG2RAD;
    sumdx = (float)cos(reg_angle0);
    sumdy = (float)sin(reg_angle0);
  }
  const float prec_deg = (float)(prec * (180.0 / VPL_PI));
  const float GUARD = 1e-2f;
  int n = 1, i = 0;
  __syncwarp();

  const int g = lane >> 3;                  // which FIFO pixel of this step (0..3)
  const int k8 = lane & 7;                  // its 8 neighbours in (yy outer, xx inner) order,
  const int k = k8 < 4 ? k8 : k8 + 1;       // skipping the centre
  const int ddx = k % 3 - 1, ddy = k / 3 - 1;

  while (i < n) {
    int take = n - i;
    if (take > 4) take = 4;
    const bool act = (g < take);
    int nidx = -1, nxy = 0;
    uint32_t ab = 0xffffffffu;
    float cs = 0.f, sn = 0.f;
    if (act) {
      int j = i + g;
      int cur = (n - j <= RING) ? e.ring[j & (RING - 1)] : (int)e.reg[j];  // packed x | y << 16
      int nx = (cur & 0xffff) + ddx, ny = (cur >> 16) + ddy;
      if (nx >= 0 && nx < ws && ny >= 0 && ny < hs) {
        nidx = ny * ws + nx;
        nxy = nx | (ny << 16);
        const Pix p = pix[nidx];  // one 128-bit gather: angle, used flag and what an acceptance adds to the sums
        ab = p.ang; cs = p.cs; sn = p.sn;
      }
    }
    bool cand = (nidx >= 0) && !(ab & kUsedBit);
    const float adeg = __uint_as_float(ab);
    // resolve acceptances in lane order == sequential order
    while (true) {
      bool al = false;
      if (cand) {
        float d = fabsf(reg_deg - adeg);
        float dd = (d > 270.f) ? fabsf(d - 360.f) : d;
        if (fabsf(dd - prec_deg) > GUARD && fabsf(d - 270.f) > GUARD) al = dd < prec_deg;
        else al = aligned_rad((double)adeg * VPL_DEG2RAD, (double)reg_deg * VPL_DEG2RAD, prec);
      }
      unsigned m = __ballot_sync(0xffffffffu, al);
      if (m == 0) break;
      int f = __ffs(m) - 1;
      int accxy = __shfl_sync(0xffffffffu, nxy, f);
      if (lane == f) {
        pix[nidx].ang = ab | kUsedBit;
        e.reg[n] = (uint32_t)nxy;
        e.ring[n & (RING - 1)] = nxy;
      }
      float fcs = __shfl_sync(0xffffffffu, cs, f);
      float fsn = __shfl_sync(0xffffffffu, sn, f);
      sumdx += fcs;
      sumdy += fsn;
      reg_deg = fast_atan2_deg(sumdy, sumdx);
      ++n;
      // lanes up to f were tested (and rejected) with the angle valid at their
      // turn; the same pixel reached through a later FIFO entry is now used.
      cand = cand && (lane > f) && (nxy != accxy);
    }
    __syncwarp();
    i += take;
  }
  *reg_angle_out = (double)reg_deg * VPL_DEG2RAD;
  return n;
}

// ---------------------------------------------------------------------------
// region2rect + get_theta (A.5).  Sequential double sums in list order.
// ---------------------------------------------------------------------------
#ifdef VPL_ENGINE_OUTLINE_MATH
__device__ __noinline__ void sincos_cr_outlined(double a, double* s, double* c) { vpl_sincos_cr(a, s, c); }
#endif
__device__ __noinline__ void region2rect(const Eng& e, int n, double reg_angle, double prec, double p, RectCand& rec) {
  const int lane = e.lane;
  double x = 0, y = 0, sum = 0;
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    int j = base + lane;
    double wx = 0, wy = 0, wt = 0;
    if (j < n) {
      const uint32_t r = e.reg[j];
      int px = (int)(r & 0xffffu), py = (int)(r >> 16);
      wt = sqrt((double)dabc_q(e.pix[py * e.ws + px].dabc) / 4.0);
      wx = (double)px * wt;
      wy = (double)py * wt;
    }
    // additions in list order: every lane reads entry t (shared-memory broadcast)
    e.bc2[lane] = make_double2(wx, wy); e.bc[lane] = wt;
    __syncwarp();
    int cnt = min(32, n - base);
    int t = 0;
    for (; t + 4 <= cnt; t += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double2 v = e.bc2[t + u];
        x += v.x;
        y += v.y;
        sum += e.bc[t + u];
      }
    }
#pragma unroll 1
    for (; t < cnt; ++t) {
      double2 v = e.bc2[t];
      x += v.x;
      y += v.y;
      sum += e.bc[t];
    }
    __syncwarp();
  }
  x /= sum;
  y /= sum;
  // get_theta
  double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    int j = base + lane;
    double t1 = 0, t2 = 0, t3 = 0;
    if (j < n) {
      const uint32_t r = e.reg[j];
      int px = (int)(r & 0xffffu), py = (int)(r >> 16);
      double weight = sqrt((double)dabc_q(e.pix[py * e.ws + px].dabc) / 4.0);
      double dx = (double)px - x, dy = (double)py - y;
      t1 = dy * dy * weight;
      t2 = dx * dx * weight;
      t3 = dx * dy * weight;
    }
    e.bc2[lane] = make_double2(t1, t2); e.bc[lane] = t3;
    __syncwarp();
    int cnt = min(32, n - base);
    int t = 0;
    for (; t + 4 <= cnt; t += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double2 v = e.bc2[t + u];
        Ixx += v.x;
        Iyy += v.y;
        Ixy -= e.bc[t + u];
      }
    }
#pragma unroll 1
    for (; t < cnt; ++t) {
      double2 v = e.bc2[t];
      Ixx += v.x;
      Iyy += v.y;
      Ixy -= e.bc[t];
    }
    __syncwarp();
  }
  double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg((float)(lambda - Ixx), (float)Ixy)
                                         : (double)fast_atan2_deg((float)Ixy, (float)(lambda - Iyy));
  theta *= VPL_DEG2RAD;
  if (fabs(angle_diff_signed_d(theta, reg_angle)) > prec) theta += VPL_PI;
  // dx = cos(theta), dy = sin(theta) through the deterministic correctly rounded sincos:
  // rect_nfa's row limits sit within an ulp of integers (the rectangle edges pass through
  // the centres of its extreme pixels), so the last bit of dx, dy decides pixel membership.
  double dx, dy;
#ifdef VPL_ENGINE_OUTLINE_MATH  // experiment for the next round (build with VPL_EXTRA_NVCC=-DVPL_ENGINE_OUTLINE_MATH): the
  sincos_cr_outlined(theta, &dy, &dx);  // same function out of line, to shrink region2rect's 40 KB of straight-line SASS
#else
  vpl_sincos_cr(theta, &dy, &dx);
#endif
  // length / width: min and max are order-free
  double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
  // (not unrolled on purpose, like the tails above: ptxas made 850 instructions of this loop, and the function is
  // straight-line code that every call streams through the instruction cache, next to 63 other warps elsewhere in the kernel)
#pragma unroll 1
  for (int j = lane; j < n; j += 32) {
    const uint32_t r = e.reg[j];
    int px = (int)(r & 0xffffu), py = (int)(r >> 16);
    double regdx = (double)px - x, regdy = (double)py - y;
    double l = regdx * dx + regdy * dy;
    double w = regdy * dx - regdx * dy;
    l_max = fmax(l_max, l); l_min = fmin(l_min, l);
    w_max = fmax(w_max, w); w_min = fmin(w_min, w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l_max = fmax(l_max, __shfl_xor_sync(0xffffffffu, l_max, o));
    l_min = fmin(l_min, __shfl_xor_sync(0xffffffffu, l_min, o));
    w_max = fmax(w_max, __shfl_xor_sync(0xffffffffu, w_max, o));
    w_min = fmin(w_min, __shfl_xor_sync(0xffffffffu, w_min, o));
  }
  rec.x1 = x + l_min * dx;
  rec.y1 = y + l_min * dy;
  rec.x2 = x + l_max * dx;
  rec.y2 = y + l_max * dy;
  rec.width = w_max - w_min;
  rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy;
  rec.prec = prec; rec.p = p;
  if (rec.width < 1.0) rec.width = 1.0;
}

// ---------------------------------------------------------------------------
// reduce_region_radius (A.6): "swap with last, pop, re-test" removal of the points
// farther than the radius.  Its result is: kept points stay in place; the k-th
// hole (ascending) among the first n_in positions receives the k-th kept point
// counted from the end.  Done as three lane-parallel passes.
// ---------------------------------------------------------------------------
__device__ __noinline__ int compact_radius(const Eng& e, int n, double xc, double yc, double radSq) {
  const int lane = e.lane;
  const unsigned lt = (1u << lane) - 1u;
  int n_in = 0;
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    int j = base + lane;
    bool in = false;
    if (j < n) {
      const uint32_t r = e.reg[j];
      int px = (int)(r & 0xffffu), py = (int)(r >> 16);
      in = !(dist_sq_d(xc, yc, (double)px, (double)py) > radSq);
      if (!in) e.pix[py * e.ws + px].ang &= ~kUsedBit;
    }
    n_in += __popc(__ballot_sync(0xffffffffu, in));
  }
  if (n_in == n) return n;
  // fillers: kept points at positions >= n_in, from the end; k-th goes to n-1-k
  int kf = 0;
  for (int top = n; top > n_in; top -= 32) {
    int j = top - 1 - lane;
    bool in = false;
    uint32_t r = 0;
    if (j >= n_in) {
      r = e.reg[j];
      int px = (int)(r & 0xffffu), py = (int)(r >> 16);
      in = !(dist_sq_d(xc, yc, (double)px, (double)py) > radSq);
    }
    unsigned m = __ballot_sync(0xffffffffu, in);
    int rank = kf + __popc(m & lt);
    __syncwarp();
    if (in) e.reg[n - 1 - rank] = r;
    kf += __popc(m);
    __syncwarp();
  }
  // holes among the first n_in positions, ascending
  int kh = 0;
#pragma unroll 1
  for (int base = 0; base < n_in; base += 32) {
    int j = base + lane;
    bool hole = false;
    if (j < n_in) {
      const uint32_t r = e.reg[j];
      int px = (int)(r & 0xffffu), py = (int)(r >> 16);
      hole = dist_sq_d(xc, yc, (double)px, (double)py) > radSq;
    }
    unsigned m = __ballot_sync(0xffffffffu, hole);
    int rank = kh + __popc(m & lt);
    if (hole) e.reg[j] = e.reg[n - 1 - rank];
    kh += __popc(m);
  }
  __syncwarp();
  return n_in;
}

__device__ bool reduce_region_radius(const Eng& e, int& n, double reg_angle, double prec, double p, RectCand& rec,
                                     double density, double density_th) {
  uint32_t s0 = e.reg[0];
  int xc_i = (int)(s0 & 0xffffu), yc_i = (int)(s0 >> 16);
  double xc = (double)xc_i, yc = (double)yc_i;
  double radSq1 = dist_sq_d(xc, yc, rec.x1, rec.y1);
  double radSq2 = dist_sq_d(xc, yc, rec.x2, rec.y2);
  double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
  while (density < density_th) {
    radSq *= 0.75 * 0.75;
    n = compact_radius(e, n, xc, yc, radSq);
    if (n < 2) return false;
    region2rect(e, n, reg_angle, prec, p, rec);
    density = (double)n / (dist_d(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  }
  return true;
}

__device__ __noinline__ bool refine(const Eng& e, int& n, double reg_angle, double prec, double p, RectCand& rec,
                       double density_th) {
  const int lane = e.lane;
  double density = (double)n / (dist_d(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  if (density >= density_th) return true;
  const uint32_t r0 = e.reg[0];
  int xc_i = (int)(r0 & 0xffffu), yc_i = (int)(r0 >> 16);
  double xc = (double)xc_i, yc = (double)yc_i;
  const float seed_deg = __uint_as_float(e.pix[yc_i * e.ws + xc_i].ang & ~kUsedBit);
  double ang_c = (double)seed_deg * VPL_DEG2RAD;
  double sum = 0, s_sum = 0;
  int cnt = 0;
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    int j = base + lane;
    bool flag = false;
    double ang_d = 0, sq = 0;
    if (j < n) {
      const uint32_t r = e.reg[j];
      int px = (int)(r & 0xffffu), py = (int)(r >> 16);
      const uint32_t au = e.pix[py * e.ws + px].ang & ~kUsedBit;
      e.pix[py * e.ws + px].ang = au;
      if (dist_d(xc, yc, (double)px, (double)py) < rec.width) {
        flag = true;
        ang_d = angle_diff_signed_d((double)__uint_as_float(au) * VPL_DEG2RAD, ang_c);
        sq = ang_d * ang_d;
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, flag);
    while (m) {
      int t = __ffs(m) - 1;
      m &= m - 1;
      sum += shfl_d(ang_d, t);
      s_sum += shfl_d(sq, t);
      ++cnt;
    }
  }
  __syncwarp();
  double mean_angle = sum / (double)cnt;
  double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)cnt + mean_angle * mean_angle);
  n = region_grow(e, (int)r0, seed_deg, tau, &reg_angle);
  if (n < 2) return false;
  region2rect(e, n, reg_angle, prec, p, rec);
  density = (double)n / (dist_d(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  if (density < density_th) return reduce_region_radius(e, n, reg_angle, prec, p, rec, density, density_th);
  return true;
}

// ---------------------------------------------------------------------------
// The engine kernel: blockDim = 32 (one warp), grid = (batch, num_octaves).
// ---------------------------------------------------------------------------
// WPB warps (= frames) per block; the warps of a block are independent.  An SM holds at most 32 blocks, so more than 32
// resident engine warps need multi-warp blocks and fewer registers per thread (MINB = blocks per SM asked of ptxas).
template <int WPB>
__device__ __forceinline__ void region_engine_body(const EngineArgs& A) {
  __shared__ int s_ring_[WPB][RING];
  __shared__ double s_bc_[WPB][32];
  __shared__ double2 s_bc2_[WPB][32];
  const int wib = threadIdx.x >> 5;
  int* s_ring = s_ring_[wib];
  double* s_bc = s_bc_[wib];
  double2* s_bc2 = s_bc2_[wib];
  const int f = blockIdx.x * WPB + wib;
  if (f >= A.batch) return;
  const EngineOct& O = A.oct[blockIdx.y];
  const size_t npx = (size_t)O.ws * O.hs;
  Eng e;
  e.pix = O.pix + (size_t)f * npx;
  e.reg = O.reg + (size_t)f * npx;
  e.ring = s_ring;
  e.bc = s_bc;
  e.bc2 = s_bc2;
  e.ws = O.ws; e.hs = O.hs;
  e.lane = threadIdx.x & 31;
  const int lane = e.lane;
  const int* ord = O.ord + (size_t)f * npx;
  const int n_ord = O.n_ord[f];
  RectCand* cand = O.cand + (size_t)f * A.cand_cap;
  const double prec = A.lc.prec, p = A.lc.p;
  const double DENSITY_TH = 0.7;
  int n_cand = 0;

  for (int base = 0; base < n_ord; base += 32) {
    int my = (base + lane < n_ord) ? ord[base + lane] : -1;
    bool free_ = false;
    if (my >= 0) free_ = !(e.pix[my].ang & kUsedBit);
    unsigned todo = __ballot_sync(0xffffffffu, free_);
    while (todo) {
      int l = __ffs(todo) - 1;
      todo &= todo - 1;
      int seed = __shfl_sync(0xffffffffu, my, l);
      // the seed may have been absorbed by a region grown earlier in this chunk
      const uint32_t sa = e.pix[seed].ang;
      if (sa & kUsedBit) continue;
      const int sy = seed / e.ws, sx = seed - sy * e.ws;
      double reg_angle;
      int n = region_grow(e, sx | (sy << 16), __uint_as_float(sa), prec, &reg_angle);
      if (n < O.min_reg_size) continue;
      RectCand rec;
      region2rect(e, n, reg_angle, prec, p, rec);
      if (!refine(e, n, reg_angle, prec, p, rec, DENSITY_TH)) continue;
      if (n_cand < A.cand_cap) {
        if (lane == 0) {
          rec.nfa = -1.0; rec.accepted = 0; rec.pad = 0;
          cand[n_cand] = rec;
        }
      } else if (lane == 0) {
        *A.overflow = 1;
      }
      ++n_cand;
    }
  }
  if (lane == 0) O.n_cand[f] = n_cand < A.cand_cap ? n_cand : A.cand_cap;
}

// Measured on B200 (C2, frames/s kernels-only at 4096 / 8192 frames per batch, same box, profiles/r02_engine_occupancy_variants.txt):
//   1 frame/block,  32 blocks/SM (64 regs, 32 warps/SM)   the round-1 shape
//   1 frame/block, no register cap (94 regs, 21 warps/SM) 44.4 k / 46.3 k
//   2 frames/block, 21 blocks/SM (40 regs, 50 warps/SM)    49.5 k / 51.0 k   <- launches of up to 3700 warps
//   2 frames/block, 25 blocks/SM (32 regs, 64 warps/SM)    51.8 k / 53.0 k   <- larger launches
//   4 frames/block, 12 blocks/SM (40 regs, 48 warps/SM)    50.2 k / 52.5 k
//   4 frames/block, 10 blocks/SM (48 regs, 40 warps/SM)    48.5 k / 52.0 k
// The engine is bound by the latency of one frame's sequential growth, so what counts is how many frames an SM holds: at
// 32 registers all 64 warp slots of an SM can be engine warps (148 x 64 = 9472 frames in flight on the device).  Two
// 4736-frame batches submitted back to back (their engine launches run side by side) reach 52.1 k frames/s end to end,
// against 47.9 k for the 40-register build, whose 50 warps per SM cannot hold both (profiles/r02_engine_wave_runs.txt).
__global__ void __launch_bounds__(32, 32) region_engine_kernel(EngineArgs A) { region_engine_body<1>(A); }
__global__ void __launch_bounds__(64, 21) region_engine_kernel_w2(EngineArgs A) { region_engine_body<2>(A); }
__global__ void __launch_bounds__(64, 25) region_engine_kernel_w2b(EngineArgs A) { region_engine_body<2>(A); }
__global__ void __launch_bounds__(128, 12) region_engine_kernel_w4(EngineArgs A) { region_engine_body<4>(A); }
__global__ void __launch_bounds__(128, 10) region_engine_kernel_w4b(EngineArgs A) { region_engine_body<4>(A); }

void launch_region_engine(const EngineArgs& a, cudaStream_t st) {
  // VPL_ENGINE_VARIANT selects one of the builds above for measurement (0: 1 frame/block, 1: 40 registers, 2: 32 registers,
  // 3, 4: four frames per block); unset, large launches take the 32-register build and small ones the 40-register one
  // (fewer spills; two such launches side by side still fit the 50 warps per SM it can hold)
  static const int forced = [] { const char* v = getenv("VPL_ENGINE_VARIANT"); return v ? atoi(v) : -1; }();
  const int variant = forced >= 0 ? forced : (a.batch * a.num_octaves > 3700 ? 2 : 1);
  dim3 grid(a.batch, a.num_octaves);
  if (variant == 1) {
    grid.x = (a.batch + 1) / 2;
    region_engine_kernel_w2<<<grid, 64, 0, st>>>(a);
  } else if (variant == 2) {
    grid.x = (a.batch + 1) / 2;
    region_engine_kernel_w2b<<<grid, 64, 0, st>>>(a);
  } else if (variant == 3) {
    grid.x = (a.batch + 3) / 4;
    region_engine_kernel_w4<<<grid, 128, 0, st>>>(a);
  } else if (variant == 4) {
    grid.x = (a.batch + 3) / 4;
    region_engine_kernel_w4b<<<grid, 128, 0, st>>>(a);
  } else {
    region_engine_kernel<<<grid, 32, 0, st>>>(a);
  }
}

}  // namespace vpl
