// lsd_engine_spec.cu -- LSD region engine, SPECULATIVE variant (sm_100a): region growing, region -> rectangle,
// refine (density test, re-grow with tighter tolerance, radius reduction) with 32 seeds of a frame in flight.
// Opt-in (vpl_debug_set_engine(ctx, 1)): bit-identical to lsd_engine.cu and to the CPU algorithm, but -- measured
// on B200, DESIGN.md section 5 -- slower than the warp-cooperative engine of lsd_engine.cu, which stays the default.
//
// Restates the main loop of cv::LineSegmentDetectorImpl::flsd and region_grow /
// region2rect / get_theta / refine / reduce_region_radius (opencv imgproc lsd.cpp;
// SURVEY.md Appendix A.4-A.6; CPU restatement oracle/orc_lsd.c).  rect_improve /
// rect_nfa do not feed back into the `used` map, so they run afterwards in a
// separate, fully parallel kernel (lsd_nfa.cu) on the candidates emitted here.
//
// Execution model: ONE WARP PER (frame, octave), 32 SPECULATIVE SEEDS IN FLIGHT.
// The sequential algorithm visits the seeds in pseudo-order and a pixel belongs to the
// first region that reaches it.  Here every lane grows the region of one seed (a
// "transaction") exactly as the sequential code would -- FIFO order, float32 running
// angle, neighbours in (yy, xx) order -- against a per-pixel owner tag:
//
//   tag[p] = 0              pixel undefined (never a candidate)
//          = 0xFFFFFFFF     free
//          = rank           owned by the transaction of seed position rank-1
//
// and the results are COMMITTED IN SEED ORDER by a commit pointer `h` (all positions
// below h are decided; ranks <= h are final).  A transaction of rank r treats a pixel as
//   used       if its tag is a committed rank, or r itself;
//   free       if the tag is FREE or a later rank (> r): taking it is a STEAL, the victim is
//              undone at once and re-run later;
//   "used, as long as the owner keeps it"  if the tag is an earlier, uncommitted rank j:
//              r records a dependency on j and is undone if j is undone or releases pixels
//              (refine / radius reduction), otherwise the sequential outcome is the same.
// Only tests of pixels whose angle is aligned matter (an unaligned pixel is rejected
// whatever its owner).  A transaction that released pixels is re-validated when it commits
// (none of the pixels it ever accepted may have ended up with an earlier rank).  With these
// rules the committed state after seed i equals the sequential state after seed i, for every
// interleaving: results are bit-identical to the CPU algorithm (oracle/orc_lsd.c), which
// tests/ assert.  A lockstep CPU model of exactly this protocol (32 and 64 lanes, tiny ring
// capacities to force every fallback) was checked against the oracle on the reference's 15
// EuRoC frames and on synthetic frames before this kernel was written.
//
// Finished transactions are PARKED (their pixel lists stay in the lane's ring of the
// frame's arena) until the commit pointer reaches them, so a lane moves on to the next free
// seed at once.  Everything that is not order-dependent inside a transaction is done by the
// whole warp: rectangle sums (per-entry products in parallel, additions in list order),
// min/max projections, releases, undo walks, commit validation.
//
// Compile with -fmad=false: the double/float expressions below mirror the CPU
// sequence operation by operation and must not be contracted.
#include "vpl_common.cuh"
#include "vpl_sincos.cuh"

#include <algorithm>

namespace vpl {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t kFree = 0xFFFFFFFFu;
constexpr int kWin = 4096;     // run-ahead window (seed positions past the commit pointer)
constexpr int kWl = 192;       // abort work list (positions)
constexpr int kNoPos = 0x7fffffff;
enum { ST_IDLE = 0, ST_GROW = 1 };
enum { DF_CAND = 1, DF_VAL = 2, DF_DEAD = 4 };

// Warp-uniform engine context (every lane holds the same values).
struct Eng {
  uint32_t* tag;      // owner tag per pixel
  const float* ang;   // level-line angle in degrees per pixel (read-only)
  const Pix* pix;     // engine records: (cosf, sinf) and the packed gradient differences of a pixel
  uint32_t* arena;    // 2*ws*hs list entries (x | y << 16)
  EngDesc* desc;      // 32 lanes x kEngQ parked descriptors
  RectCand* rects;    // 32 lanes x kEngQ staged rectangles
  const int* ord;
  int n_ord;
  int ws, hs;
  int capl;           // ring capacity per lane (power of two)
  int arena_n;        // entries in the arena
  int min_reg;
  double prec, p;
  // shared memory
  int* wl;            // abort work list
  int* steals;        // 32 x 8 victim positions
  double* bc;         // 32 doubles
  double2* bc2;       // 32 double2
  int lane;
};

// Per-lane state (registers).
struct LaneSt {
  int st, pos;
  int rstart;         // absolute ring index where the current transaction's lists start
  int front;          // absolute ring index of the oldest live parked entry (== rstart when none)
  int base, n, i, kst, phase, ext;
  float reg_deg, sumdx, sumdy, prec_deg;
  double prec;
  int nd, dep0, dep1, dep2, dep3;
  int need_val;
  int qh, qn;         // parked queue (circular, kEngQ slots, includes dead entries)
  int big;            // this lane runs in big mode (whole arena, no wrap)
};

// index of the n-th (0-based) set bit of m
__device__ __forceinline__ int nth_set(unsigned m, int n) {
  for (int i = 0; i < n; ++i) m &= m - 1;
  return __ffs(m) - 1;
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }
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
// The running region angle is always (double)reg_deg * DEG2RAD with reg_deg the float32 fastAtan2
// result, so the alignment test is first decided in float32 on the degrees and only falls back to
// the exact double sequence inside a guard band around the thresholds (float error on the
// difference is < 1e-4 deg; the band is 1e-2 deg): same decisions, a fraction of the FP64 work.
__device__ __forceinline__ bool aligned_deg(float adeg, float reg_deg, float prec_deg, double prec) {
  const float GUARD = 1e-2f;
  float d = fabsf(reg_deg - adeg);
  float dd = (d > 270.f) ? fabsf(d - 360.f) : d;
  if (fabsf(dd - prec_deg) > GUARD && fabsf(d - 270.f) > GUARD) return dd < prec_deg;
  return aligned_rad((double)adeg * VPL_DEG2RAD, (double)reg_deg * VPL_DEG2RAD, prec);
}

// A transaction's list storage, warp-uniform: entry j lives at base[(start + j) & mask].
struct ListRef {
  uint32_t* base;
  int start;
  int mask;
  __device__ __forceinline__ uint32_t& at(int j) const { return base[(start + j) & mask]; }
};
__device__ __forceinline__ ListRef list_of(const Eng& e, int owner_lane, int start, int big) {
  ListRef r;
  r.base = big ? e.arena : e.arena + (size_t)owner_lane * e.capl;
  r.start = start;
  r.mask = big ? 0x7fffffff : e.capl - 1;
  return r;
}

// ---------------------------------------------------------------------------
// region2rect + get_theta (A.5), warp-cooperative.  Sequential double sums in list order:
// per-entry products in parallel, the additions themselves in list order through a
// shared-memory broadcast, so the results equal the sequential sums bit for bit.
// ---------------------------------------------------------------------------
__device__ __noinline__ void region2rect(const Eng& e, const ListRef& L, int off, int n, double reg_angle, double prec,
                                         double p, RectCand& rec) {
  const int lane = e.lane;
  double x = 0, y = 0, sum = 0;
  for (int b0 = 0; b0 < n; b0 += 32) {
    int j = b0 + lane;
    double wx = 0, wy = 0, wt = 0;
    if (j < n) {
      const uint32_t r = L.at(off + j);
      int px = (int)(r & 0xffffu), py = (int)(r >> 16);
      wt = sqrt((double)dabc_q(e.pix[py * e.ws + px].dabc) / 4.0);
      wx = (double)px * wt;
      wy = (double)py * wt;
    }
    e.bc2[lane] = make_double2(wx, wy); e.bc[lane] = wt;
    __syncwarp();
    int cnt = min(32, n - b0);
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
  double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
  for (int b0 = 0; b0 < n; b0 += 32) {
    int j = b0 + lane;
    double t1 = 0, t2 = 0, t3 = 0;
    if (j < n) {
      const uint32_t r = L.at(off + j);
      int px = (int)(r & 0xffffu), py = (int)(r >> 16);
      double weight = sqrt((double)dabc_q(e.pix[py * e.ws + px].dabc) / 4.0);
      double dx = (double)px - x, dy = (double)py - y;
      t1 = dy * dy * weight;
      t2 = dx * dx * weight;
      t3 = dx * dy * weight;
    }
    e.bc2[lane] = make_double2(t1, t2); e.bc[lane] = t3;
    __syncwarp();
    int cnt = min(32, n - b0);
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
  vpl_sincos_cr(theta, &dy, &dx);
  double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
  for (int j = lane; j < n; j += 32) {
    const uint32_t r = L.at(off + j);
    int px = (int)(r & 0xffffu), py = (int)(r >> 16);
    double regdx = (double)px - x, regdy = (double)py - y;
    double l = regdx * dx + regdy * dy;
    double w = regdy * dx - regdx * dy;
    l_max = fmax(l_max, l); l_min = fmin(l_min, l);
    w_max = fmax(w_max, w); w_min = fmin(w_min, w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l_max = fmax(l_max, __shfl_xor_sync(FULL, l_max, o));
    l_min = fmin(l_min, __shfl_xor_sync(FULL, l_min, o));
    w_max = fmax(w_max, __shfl_xor_sync(FULL, w_max, o));
    w_min = fmin(w_min, __shfl_xor_sync(FULL, w_min, o));
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
// reduce_region_radius's removal pass (A.6): "swap with last, pop, re-test" of the points
// farther than the radius.  Its result is: kept points stay in place; the k-th hole
// (ascending) among the first n_in positions receives the k-th kept point counted from the
// end.  Done as three lane-parallel passes.  Removed pixels are released (tag FREE).
// ---------------------------------------------------------------------------
__device__ __noinline__ int compact_radius(const Eng& e, const ListRef& L, int off, int n, uint32_t rank, double xc,
                                           double yc, double radSq) {
  const int lane = e.lane;
  const unsigned lt = (1u << lane) - 1u;
  int n_in = 0;
  for (int b0 = 0; b0 < n; b0 += 32) {
    int j = b0 + lane;
    bool in = false;
    if (j < n) {
      const uint32_t r = L.at(off + j);
      int px = (int)(r & 0xffffu), py = (int)(r >> 16);
      in = !(dist_sq_d(xc, yc, (double)px, (double)py) > radSq);
      if (!in) {
        uint32_t* t = e.tag + (size_t)py * e.ws + px;
        if (*t == rank) *t = kFree;
      }
    }
    n_in += __popc(__ballot_sync(FULL, in));
  }
  if (n_in == n) return n;
  int kf = 0;
  for (int top = n; top > n_in; top -= 32) {
    int j = top - 1 - lane;
    bool in = false;
    uint32_t r = 0;
    if (j >= n_in) {
      r = L.at(off + j);
      int px = (int)(r & 0xffffu), py = (int)(r >> 16);
      in = !(dist_sq_d(xc, yc, (double)px, (double)py) > radSq);
    }
    unsigned m = __ballot_sync(FULL, in);
    int rk = kf + __popc(m & lt);
    __syncwarp();
    if (in) L.at(off + n - 1 - rk) = r;
    kf += __popc(m);
    __syncwarp();
  }
  int kh = 0;
  for (int b0 = 0; b0 < n_in; b0 += 32) {
    int j = b0 + lane;
    bool hole = false;
    if (j < n_in) {
      const uint32_t r = L.at(off + j);
      int px = (int)(r & 0xffffu), py = (int)(r >> 16);
      hole = dist_sq_d(xc, yc, (double)px, (double)py) > radSq;
    }
    unsigned m = __ballot_sync(FULL, hole);
    int rk = kh + __popc(m & lt);
    if (hole) L.at(off + j) = L.at(off + n - 1 - rk);
    kh += __popc(m);
  }
  __syncwarp();
  return n_in;
}

// ---------------------------------------------------------------------------
// Parked-queue helpers (lane-private data in global memory: e.desc[lane * kEngQ + slot]).
// ---------------------------------------------------------------------------
__device__ __forceinline__ EngDesc* dslot(const Eng& e, int slot) { return e.desc + e.lane * kEngQ + slot; }

// drop dead entries from the front of the lane's queue and recompute `front`
__device__ __forceinline__ void pop_dead_front(const Eng& e, LaneSt& s) {
  while (s.qn > 0 && (dslot(e, s.qh)->flags & DF_DEAD)) {
    s.qh = (s.qh + 1) & (kEngQ - 1);
    --s.qn;
  }
  if (s.qn > 0) s.front = dslot(e, s.qh)->start;
  else if (s.st == ST_IDLE) { s.rstart = 0; s.front = 0; }  // nothing live: restart the ring
  else s.front = s.rstart;
}

// slot of the live parked transaction of position v in this lane's queue, or -1
__device__ __forceinline__ int find_parked(const Eng& e, const LaneSt& s, int v) {
  for (int k = 0; k < s.qn; ++k) {
    int sl = (s.qh + k) & (kEngQ - 1);
    const EngDesc* d = dslot(e, sl);
    if (d->pos == v && !(d->flags & DF_DEAD)) return sl;
  }
  return -1;
}

__device__ __forceinline__ bool cur_depends(const LaneSt& s, int v) {
  return s.st != ST_IDLE && ((s.nd > 0 && s.dep0 == v) || (s.nd > 1 && s.dep1 == v) || (s.nd > 2 && s.dep2 == v) ||
                             (s.nd > 3 && s.dep3 == v));
}

// ---------------------------------------------------------------------------
// Abort machinery (warp-cooperative).  wl[0..wn) holds positions; bit 30 set = "released
// pixels" event (no undo, only the dependants are aborted).
// ---------------------------------------------------------------------------
constexpr int kRelFlag = 1 << 30;

struct Ctl {           // warp-uniform control registers
  int h;               // commit pointer
  int nxt;             // forward scan pointer
  int rp;              // retry scan pointer (kNoPos = nothing to retry)
  int wn;              // entries in the work list
  int n_cand;
  int bigmode;
  int fail;            // work list overflow etc.: fall back to aborting everything speculative
};

__device__ __forceinline__ void wl_push_uniform(const Eng& e, Ctl& c, int v) {
  if (c.wn < kWl) {
    if (e.lane == 0) e.wl[c.wn] = v;
    ++c.wn;
  } else {
    c.fail = 1;
  }
}

// push from the lanes whose `want` is set (value `v` per lane)
__device__ __forceinline__ void wl_push_lanes(const Eng& e, Ctl& c, bool want, int v) {
  unsigned m = __ballot_sync(FULL, want);
  if (!m) return;
  int cnt = __popc(m);
  if (c.wn + cnt > kWl) { c.fail = 1; return; }
  if (want) e.wl[c.wn + __popc(m & ((1u << e.lane) - 1u))] = v;
  c.wn += cnt;
}

// undo every pixel the transaction of position v still owns, cooperative
__device__ __forceinline__ void undo_walk(const Eng& e, const ListRef& L, int ext, uint32_t rank) {
  for (int j = e.lane; j < ext; j += 32) {
    const uint32_t r = L.at(j);
    uint32_t* t = e.tag + (size_t)(r >> 16) * e.ws + (r & 0xffffu);
    if (*t == rank) *t = kFree;
  }
}

// every lane pushes its transactions (current, parked) that depend on position v
__device__ __forceinline__ void push_dependants(const Eng& e, Ctl& c, LaneSt& s, int v) {
  wl_push_lanes(e, c, cur_depends(s, v) && s.pos > v, s.pos);
  unsigned any = __ballot_sync(FULL, s.qn > 0);
  if (!any) return;
  int maxq = s.qn;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxq = max(maxq, __shfl_xor_sync(FULL, maxq, o));
  for (int k = 0; k < maxq; ++k) {
    bool want = false;
    int pv = 0;
    if (k < s.qn) {
      const EngDesc* d = dslot(e, (s.qh + k) & (kEngQ - 1));
      if (!(d->flags & DF_DEAD) && d->pos > v) {
        int nd = d->nd;
        want = (nd > 0 && d->dep[0] == v) || (nd > 1 && d->dep[1] == v) || (nd > 2 && d->dep[2] == v) ||
               (nd > 3 && d->dep[3] == v);
        pv = d->pos;
      }
    }
    wl_push_lanes(e, c, want, pv);
  }
}

// process the work list until it is empty
__device__ __noinline__ void process_aborts(const Eng& e, Ctl& c, LaneSt& s) {
  int done = 0;
  while (done < c.wn) {
    __syncwarp();
    int item = e.wl[done++];
    int v = item & ~kRelFlag;
    if (!(item & kRelFlag)) {
      // locate the holder
      bool cur = (s.st != ST_IDLE && s.pos == v);
      int sl = cur ? -1 : find_parked(e, s, v);
      unsigned mh = __ballot_sync(FULL, cur || sl >= 0);
      if (!mh) continue;  // already undone
      int owner = __ffs(mh) - 1;
      int start = cur ? s.rstart : (sl >= 0 ? dslot(e, sl)->start : 0);
      int ext = cur ? s.ext : (sl >= 0 ? dslot(e, sl)->ext : 0);
      start = __shfl_sync(FULL, start, owner);
      ext = __shfl_sync(FULL, ext, owner);
      int big = __shfl_sync(FULL, s.big, owner);
      ListRef L = list_of(e, owner, start, big);
      undo_walk(e, L, ext, (uint32_t)v + 1u);
      __syncwarp();
      if (e.lane == owner) {
        if (cur) { s.st = ST_IDLE; s.nd = 0; }
        else dslot(e, sl)->flags |= DF_DEAD;
        pop_dead_front(e, s);
      }
      if (v < c.rp) c.rp = v;
    }
    push_dependants(e, c, s, v);
    if (c.fail) break;
  }
  c.wn = 0;
  __syncwarp();
}

// Fallback (work list overflow): undo every transaction except the one at the commit pointer.
__device__ __noinline__ void abort_all_speculative(const Eng& e, Ctl& c, LaneSt& s) {
  c.fail = 0; c.wn = 0;
  for (int owner = 0; owner < 32; ++owner) {
    // current
    int st = __shfl_sync(FULL, s.st, owner), pos = __shfl_sync(FULL, s.pos, owner);
    if (st != ST_IDLE && pos != c.h) {
      ListRef L = list_of(e, owner, __shfl_sync(FULL, s.rstart, owner), __shfl_sync(FULL, s.big, owner));
      undo_walk(e, L, __shfl_sync(FULL, s.ext, owner), (uint32_t)pos + 1u);
      if (e.lane == owner) { s.st = ST_IDLE; s.nd = 0; }
    }
    int qn = __shfl_sync(FULL, s.qn, owner), qh = __shfl_sync(FULL, s.qh, owner);
    for (int k = 0; k < qn; ++k) {
      EngDesc* d = e.desc + owner * kEngQ + ((qh + k) & (kEngQ - 1));
      int fl = d->flags, dpos = d->pos, dstart = d->start, dext = d->ext;
      if (fl & DF_DEAD) continue;
      ListRef L = list_of(e, owner, dstart, 0);
      undo_walk(e, L, dext, (uint32_t)dpos + 1u);
      __syncwarp();
      if (e.lane == owner) d->flags = fl | DF_DEAD;
    }
    __syncwarp();
    if (e.lane == owner) pop_dead_front(e, s);
  }
  c.rp = c.h;
  __syncwarp();
}

// ---------------------------------------------------------------------------
// Start the transaction of seed position `pos` on this lane (lane-private; the seed pixel
// must be free -- the caller has undone a later owner if there was one).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void lane_begin(const Eng& e, LaneSt& s, int pos) {
  const int seed = e.ord[pos];
  const int sy = seed / e.ws, sx = seed - sy * e.ws;
  const float seed_deg = e.ang[seed];
  e.tag[seed] = (uint32_t)pos + 1u;
  uint32_t* base = s.big ? e.arena : e.arena + (size_t)e.lane * e.capl;
  base[s.rstart & (s.big ? 0x7fffffff : e.capl - 1)] = (uint32_t)sx | ((uint32_t)sy << 16);
  s.st = ST_GROW; s.pos = pos; s.base = 0; s.n = 1; s.i = 0; s.kst = 0; s.phase = 0; s.ext = 1;
  s.nd = 0; s.need_val = 0;
  s.reg_deg = seed_deg;
  const double a0 = (double)seed_deg * VPL_DEG2RAD;
  s.sumdx = (float)cos(a0);
  s.sumdy = (float)sin(a0);
  s.prec = e.prec;
  s.prec_deg = (float)(e.prec * (180.0 / VPL_PI));
}

// ---------------------------------------------------------------------------
// What follows the growth of lane `owner`'s current list (warp-cooperative): size gate,
// rectangle, refine (A.6).  Returns with the lane either growing again (refine's re-grow)
// or finished: *done = 1, *has_cand says whether `rec` is a candidate.
// ---------------------------------------------------------------------------
__device__ __noinline__ void after_grow(const Eng& e, Ctl& c, LaneSt& s, int owner, RectCand& rec, int* done,
                                        int* has_cand) {
  const int lane = e.lane;
  const int pos = __shfl_sync(FULL, s.pos, owner);
  const int phase = __shfl_sync(FULL, s.phase, owner);
  int n = __shfl_sync(FULL, s.n, owner);
  const int base = __shfl_sync(FULL, s.base, owner);
  const float reg_deg = __shfl_sync(FULL, s.reg_deg, owner);
  const double reg_angle = (double)reg_deg * VPL_DEG2RAD;
  const ListRef L = list_of(e, owner, __shfl_sync(FULL, s.rstart, owner), __shfl_sync(FULL, s.big, owner));
  const uint32_t rank = (uint32_t)pos + 1u;
  const double DENSITY_TH = 0.7;
  *done = 1; *has_cand = 0;
  if (phase == 0) {
    if (n < e.min_reg) return;
    region2rect(e, L, base, n, reg_angle, e.prec, e.p, rec);
    double density = (double)n / (dist_d(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    if (density >= DENSITY_TH) { *has_cand = 1; return; }
    // refine: release every pixel of the region, statistics of the angles near the seed, re-grow
    const uint32_t r0 = L.at(base);
    const int xc_i = (int)(r0 & 0xffffu), yc_i = (int)(r0 >> 16);
    const double xc = (double)xc_i, yc = (double)yc_i;
    uint32_t* t0 = e.tag + (size_t)yc_i * e.ws + xc_i;
    const float seed_deg = e.ang[(size_t)yc_i * e.ws + xc_i];
    const double ang_c = (double)seed_deg * VPL_DEG2RAD;
    double sum = 0, s_sum = 0;
    int cnt = 0;
    for (int b0 = 0; b0 < n; b0 += 32) {
      int j = b0 + lane;
      bool flag = false;
      double ang_d = 0, sq = 0;
      if (j < n) {
        const uint32_t r = L.at(base + j);
        int px = (int)(r & 0xffffu), py = (int)(r >> 16);
        uint32_t* t = e.tag + (size_t)py * e.ws + px;
        if (*t == rank) *t = kFree;
        if (dist_d(xc, yc, (double)px, (double)py) < rec.width) {
          flag = true;
          ang_d = angle_diff_signed_d((double)e.ang[(size_t)py * e.ws + px] * VPL_DEG2RAD, ang_c);
          sq = ang_d * ang_d;
        }
      }
      unsigned m = __ballot_sync(FULL, flag);
      while (m) {
        int t = __ffs(m) - 1;
        m &= m - 1;
        sum += shfl_d(ang_d, t);
        s_sum += shfl_d(sq, t);
        ++cnt;
      }
    }
    __syncwarp();
    wl_push_uniform(e, c, pos | kRelFlag);  // whoever assumed these pixels used must be re-run
    if (pos + 1 < c.rp) c.rp = pos + 1;     // seeds among the released pixels may start now
    double mean_angle = sum / (double)cnt;
    double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)cnt + mean_angle * mean_angle);
    // second list right after the first; it starts with the seed again
    if (lane == owner) {
      L.at(base + n) = r0;
      *t0 = rank;
      s.need_val = 1;
      s.base = base + n; s.phase = 1; s.n = 1; s.i = 0; s.kst = 0;
      s.ext = base + n + 1;
      s.reg_deg = seed_deg;
      s.sumdx = (float)cos(ang_c);
      s.sumdy = (float)sin(ang_c);
      s.prec = tau;
      s.prec_deg = (float)(tau * (180.0 / VPL_PI));
    }
    __syncwarp();
    *done = 0;
    return;
  }
  // after the re-grow
  if (n < 2) return;
  region2rect(e, L, base, n, reg_angle, e.prec, e.p, rec);
  double density = (double)n / (dist_d(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  if (density < DENSITY_TH) {
    const uint32_t r0 = L.at(base);
    const double xc = (double)(int)(r0 & 0xffffu), yc = (double)(int)(r0 >> 16);
    double radSq1 = dist_sq_d(xc, yc, rec.x1, rec.y1);
    double radSq2 = dist_sq_d(xc, yc, rec.x2, rec.y2);
    double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
    wl_push_uniform(e, c, pos | kRelFlag);
    if (pos + 1 < c.rp) c.rp = pos + 1;
    while (density < DENSITY_TH) {
      radSq *= 0.75 * 0.75;
      n = compact_radius(e, L, base, n, rank, xc, yc, radSq);
      if (n < 2) break;
      region2rect(e, L, base, n, reg_angle, e.prec, e.p, rec);
      density = (double)n / (dist_d(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    }
    if (lane == owner) s.n = n;
    if (n < 2) return;
  }
  *has_cand = 1;
}

// commit-time validation of a transaction that released pixels: none of the pixels it ever
// accepted may have ended up with an earlier rank (warp-cooperative) -> true if valid
__device__ __forceinline__ bool validate_walk(const Eng& e, const ListRef& L, int ext, uint32_t rank) {
  bool bad = false;
  for (int j = e.lane; j < ext; j += 32) {
    const uint32_t r = L.at(j);
    const uint32_t t = e.tag[(size_t)(r >> 16) * e.ws + (r & 0xffffu)];
    bad |= t < rank;
  }
  return __ballot_sync(FULL, bad) == 0;
}

__device__ __forceinline__ void emit_cand(const Eng& e, Ctl& c, RectCand* cand, int cand_cap, int* overflow,
                                          const RectCand& rec) {
  if (c.n_cand < cand_cap) {
    if (e.lane == 0) {
      RectCand r = rec;
      r.nfa = -1.0; r.accepted = 0; r.pad = 0;
      cand[c.n_cand] = r;
    }
  } else if (e.lane == 0) {
    *overflow = 1;
  }
  ++c.n_cand;
}

}  // namespace

// ---------------------------------------------------------------------------
// The engine kernel: blockDim = 32 (one warp), grid = (batch, num_octaves).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
region_engine_spec_kernel(EngineArgs A) {
  __shared__ int s_wl[kWl];
  __shared__ int s_steals[32 * 8];
  __shared__ double s_bc[32];
  __shared__ double2 s_bc2[32];
  const int f = blockIdx.x;
  const EngineOct& O = A.oct[blockIdx.y];
  const size_t npx = (size_t)O.ws * O.hs;
  Eng e;
  e.tag = O.tag + (size_t)f * npx;
  e.ang = O.ang + (size_t)f * npx;
  e.pix = O.pix + (size_t)f * npx;
  e.arena = O.arena + (size_t)f * 2 * npx;
  e.desc = O.desc + (size_t)f * 32 * kEngQ;
  e.rects = O.rects + (size_t)f * 32 * kEngQ;
  e.ord = O.ord + (size_t)f * npx;
  e.n_ord = O.n_ord[f];
  e.ws = O.ws; e.hs = O.hs;
  e.arena_n = (int)(2 * npx);
  {
    int capl = 1;
    const int lim = A.ring_cap > 0 ? A.ring_cap : (int)(2 * npx / 32);
    while (capl * 2 <= lim) capl *= 2;
    e.capl = capl;
  }
  e.min_reg = O.min_reg_size;
  e.prec = A.lc.prec; e.p = A.lc.p;
  e.wl = s_wl; e.steals = s_steals; e.bc = s_bc; e.bc2 = s_bc2;
  e.lane = threadIdx.x;
  const int lane = e.lane;
  const unsigned lt = (1u << lane) - 1u;
  RectCand* cand = O.cand + (size_t)f * A.cand_cap;

  LaneSt s;
  s.st = ST_IDLE; s.pos = 0; s.rstart = 0; s.front = 0; s.base = 0; s.n = 0; s.i = 0; s.kst = 0; s.phase = 0; s.ext = 0;
  s.reg_deg = 0.f; s.sumdx = 0.f; s.sumdy = 0.f; s.prec_deg = 0.f; s.prec = 0.0;
  s.nd = 0; s.dep0 = s.dep1 = s.dep2 = s.dep3 = 0; s.need_val = 0; s.qh = 0; s.qn = 0; s.big = 0;
  Ctl c;
  c.h = 0; c.nxt = 0; c.rp = kNoPos; c.wn = 0; c.n_cand = 0; c.bigmode = 0; c.fail = 0;
  const int n_ord = e.n_ord;
  long long guard = 0;
  const long long guard_max = 64LL * (long long)npx + 100000;

  for (;;) {
    // ================= 1. commit sweep =================
    bool head_needs_lane = false;
    while (c.h < n_ord) {
      // leading run of dead seeds, 32 at a time
      int ppos = c.h + lane;
      uint32_t t = 0;
      if (ppos < n_ord) t = e.tag[e.ord[ppos]];
      unsigned dead = __ballot_sync(FULL, ppos < n_ord && t <= (uint32_t)c.h);  // owned by a committed rank
      int run = __ffs(~dead) - 1;  // 32 -> ffs(0) = 0 -> -1
      if (run < 0) run = 32;
      if (run > 0) { c.h += run; continue; }
      // position h is not dead
      uint32_t th = __shfl_sync(FULL, t, 0);
      if (th == (uint32_t)c.h + 1u) {
        // held: running or parked
        unsigned mr = __ballot_sync(FULL, s.st != ST_IDLE && s.pos == c.h);
        if (mr) break;  // still running
        int sl = find_parked(e, s, c.h);
        unsigned mp = __ballot_sync(FULL, sl >= 0);
        if (!mp) {  // cannot happen (a held seed has a holder); treat as free
          if (lane == 0) e.tag[e.ord[c.h]] = kFree;
          __syncwarp();
          continue;
        }
        int owner = __ffs(mp) - 1;
        int fl = 0, start = 0, ext = 0;
        if (lane == owner) { const EngDesc* d = dslot(e, sl); fl = d->flags; start = d->start; ext = d->ext; }
        fl = __shfl_sync(FULL, fl, owner); start = __shfl_sync(FULL, start, owner); ext = __shfl_sync(FULL, ext, owner);
        int osl = __shfl_sync(FULL, sl, owner);
        if (fl & DF_VAL) {
          ListRef L = list_of(e, owner, start, 0);
          if (!validate_walk(e, L, ext, (uint32_t)c.h + 1u)) {
            wl_push_uniform(e, c, c.h);
            process_aborts(e, c, s);
            if (c.fail) abort_all_speculative(e, c, s);
            continue;  // the seed is re-evaluated (free now): it runs again as the head
          }
        }
        if (fl & DF_CAND) emit_cand(e, c, cand, A.cand_cap, A.overflow, e.rects[owner * kEngQ + osl]);
        if (lane == owner) {
          dslot(e, sl)->flags = fl | DF_DEAD;
          pop_dead_front(e, s);
        }
        __syncwarp();
        ++c.h;
        continue;
      }
      head_needs_lane = true;  // free, or owned by a later (speculative) transaction
      break;
    }
    if (c.nxt < c.h) c.nxt = c.h;
    {
      unsigned busy = __ballot_sync(FULL, s.st != ST_IDLE || s.qn > 0);
      if (c.h >= n_ord && !busy) break;
    }
    if (++guard > guard_max) {  // watchdog: never hang the device
      if (lane == 0) *A.overflow = 2;
      break;
    }
    if (c.bigmode) {
      unsigned busy = __ballot_sync(FULL, s.st != ST_IDLE);
      if (!busy) {  // the big transaction has committed: back to the partitioned rings
        c.bigmode = 0;
        s.big = 0; s.rstart = 0; s.front = 0; s.qh = 0; s.qn = 0;
      }
    }

    // ================= 2. assignment =================
    if (head_needs_lane) {
      // a later owner of the seed pixel is undone first
      uint32_t th = e.tag[e.ord[c.h]];
      if (th != kFree) {
        wl_push_uniform(e, c, (int)th - 1);
        process_aborts(e, c, s);
        if (c.fail) abort_all_speculative(e, c, s);
      }
      unsigned idle = __ballot_sync(FULL, s.st == ST_IDLE);
      if (!idle) {  // preempt the running transaction with the highest position
        int mp = s.pos;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mp = max(mp, __shfl_xor_sync(FULL, mp, o));
        wl_push_uniform(e, c, mp);
        process_aborts(e, c, s);
        if (c.fail) abort_all_speculative(e, c, s);
        idle = __ballot_sync(FULL, s.st == ST_IDLE);
      }
      // the idle lane with the most free ring space takes the head
      int fr = (s.st == ST_IDLE) ? (e.capl - (s.rstart - s.front)) : -1;
      int best = fr;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
      unsigned mb = __ballot_sync(FULL, fr == best && s.st == ST_IDLE);
      int owner = __ffs(mb) - 1;
      if (best < 32) {
        // no room even for the first entries: what this lane has parked is undone (it runs again later)
        for (;;) {
          int dp = -1;
          if (lane == owner)
            for (int k = 0; k < s.qn && dp < 0; ++k) {
              const EngDesc* d = dslot(e, (s.qh + k) & (kEngQ - 1));
              if (!(d->flags & DF_DEAD)) dp = d->pos;
            }
          dp = __shfl_sync(FULL, dp, owner);
          if (dp < 0) break;
          wl_push_uniform(e, c, dp);
          process_aborts(e, c, s);
          if (c.fail) abort_all_speculative(e, c, s);
        }
        if (lane == owner) pop_dead_front(e, s);
      }
      if (lane == owner) lane_begin(e, s, c.h);
      __syncwarp();
      if (c.nxt <= c.h) c.nxt = c.h + 1;
    }
    if (!c.bigmode) {
      // lanes that can take run-ahead work: idle, a free parking slot, a quarter of the ring free
      // (two passes: retry scan over [rp, nxt), then the forward scan from nxt)
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        bool avail = s.st == ST_IDLE && s.qn < kEngQ && (e.capl - (s.rstart - s.front)) >= (e.capl >> 2);
        unsigned ma = __ballot_sync(FULL, avail);
        if (!ma) break;
        int from, lim;
        if (pass == 0) {
          if (c.rp < c.h) c.rp = c.h;
          if (c.rp >= c.nxt) { c.rp = kNoPos; continue; }
          from = c.rp; lim = c.nxt;
        } else {
          if (c.nxt >= n_ord || c.nxt - c.h >= kWin) break;
          from = c.nxt; lim = n_ord;
        }
        int ppos = from + lane;
        bool startable = false;
        if (ppos < lim) startable = e.tag[e.ord[ppos]] == kFree;
        unsigned ms = __ballot_sync(FULL, startable);
        int na = min(__popc(ms), __popc(ma));
        int my = __popc(ma & lt);
        if (avail && my < na) {
          int src = nth_set(ms, my);
          lane_begin(e, s, from + src);
        }
        int adv = 32;
        if (__popc(ms) > na) adv = nth_set(ms, na - 1) + 1;  // stop right after the last seed handed out
        if (pass == 0) {
          c.rp = from + adv;
          if (c.rp >= c.nxt) c.rp = kNoPos;
        } else {
          c.nxt = min(from + adv, n_ord);
        }
        __syncwarp();
      }
    }

    // ================= 3. ring space of the growing lanes =================
    {
      bool trouble = s.st == ST_GROW && !s.big && (s.rstart + s.ext + 9 - s.front) > e.capl;
      unsigned mt = __ballot_sync(FULL, trouble);
      while (mt) {
        int owner = __ffs(mt) - 1;
        mt &= mt - 1;
        int opos = __shfl_sync(FULL, s.pos, owner);
        if (opos != c.h) {  // a speculative transaction: undo it, it runs again later
          wl_push_uniform(e, c, opos);
          process_aborts(e, c, s);
          if (c.fail) abort_all_speculative(e, c, s);
          continue;
        }
        // the head: drop what its lane has parked, then, if that is not enough, big mode
        for (;;) {
          int sl = -1;
          if (lane == owner)
            for (int k = 0; k < s.qn && sl < 0; ++k) {
              int q = (s.qh + k) & (kEngQ - 1);
              if (!(dslot(e, q)->flags & DF_DEAD)) sl = q;
            }
          int dp = (lane == owner && sl >= 0) ? dslot(e, sl)->pos : -1;
          dp = __shfl_sync(FULL, dp, owner);
          if (dp < 0) break;
          wl_push_uniform(e, c, dp);
          process_aborts(e, c, s);
          if (c.fail) abort_all_speculative(e, c, s);
        }
        bool still = false;
        if (lane == owner) still = s.st == ST_GROW && (s.rstart + s.ext + 9 - s.front) > e.capl;
        still = __shfl_sync(FULL, (int)still, owner) != 0;
        if (still) {
          // big mode: everything else is undone, the head restarts with the whole arena
          abort_all_speculative(e, c, s);
          wl_push_uniform(e, c, c.h);
          process_aborts(e, c, s);
          c.bigmode = 1; c.rp = kNoPos; c.nxt = c.h + 1;
          if (lane == owner) {
            s.big = 1; s.rstart = 0; s.front = 0; s.qh = 0; s.qn = 0;
            lane_begin(e, s, c.h);
          }
          __syncwarp();
        }
      }
    }

    // ================= 4. one growth step: the 8 neighbours of every growing lane's FIFO pixel =================
    // (the lane state touched here is held in plain locals so that it stays in registers)
    int nsteal = 0;
    bool blocked = false;
    {
      const bool grow = s.st == ST_GROW;
      const int kst = s.kst, g_base = s.rstart + s.base;
      const uint32_t rank = (uint32_t)s.pos + 1u, hh = (uint32_t)c.h;
      const float prec_deg = s.prec_deg;
      const double prec = s.prec;
      float reg_deg = s.reg_deg, sumdx = s.sumdx, sumdy = s.sumdy;
      int n = s.n, nd = s.nd, d0 = s.dep0, d1 = s.dep1, d2 = s.dep2, d3 = s.dep3, bk = 0;
      uint2 nb[8];
      int nidx[8];
      int cx = 0, cy = 0;
      uint32_t* lbase = s.big ? e.arena : e.arena + (size_t)lane * e.capl;
      const int lmask = s.big ? 0x7fffffff : e.capl - 1;
      if (grow) {
        const uint32_t cur = lbase[(g_base + s.i) & lmask];
        cx = (int)(cur & 0xffffu); cy = (int)(cur >> 16);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int kk = k < 4 ? k : k + 1;
        const int nx = cx + (kk % 3 - 1), ny = cy + (kk / 3 - 1);
        nidx[k] = -1;
        nb[k] = make_uint2(0u, 0u);
        if (grow && k >= kst && nx >= 0 && nx < e.ws && ny >= 0 && ny < e.hs) {
          nidx[k] = ny * e.ws + nx;
          nb[k] = make_uint2(__float_as_uint(__ldg(e.ang + nidx[k])), e.tag[nidx[k]]);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        bool prop = false;
        if (!blocked && nidx[k] >= 0) {
          const uint32_t t = nb[k].y;
          // candidate unless undefined, mine, or owned by a committed rank
          if (t != 0u && t != rank && t > hh) prop = aligned_deg(__uint_as_float(nb[k].x), reg_deg, prec_deg, prec);
        }
        const unsigned mp = __ballot_sync(FULL, prop);
        if (!mp) continue;
        if (prop) {
          // the tag may have changed during this step: look again
          uint32_t* tp = e.tag + nidx[k];
          const uint32_t t = *tp;
          bool take = false;
          if (t == kFree || t > rank) take = true;
          else if (t != rank && t > hh) {
            // owned by an earlier, uncommitted transaction: assume it keeps the pixel
            const int j = (int)t - 1;
            const bool have = (nd > 0 && d0 == j) || (nd > 1 && d1 == j) || (nd > 2 && d2 == j) || (nd > 3 && d3 == j);
            if (!have) {
              if (nd == 0) d0 = j;
              else if (nd == 1) d1 = j;
              else if (nd == 2) d2 = j;
              else if (nd == 3) d3 = j;
              if (nd < kEngDeps) ++nd;
              else { blocked = true; bk = k; }  // no room for another dependency: wait for that one to commit
            }
          }
          // several lanes may want the same pixel in this step: the lowest rank takes it
          const unsigned peers = __match_any_sync(mp, take ? nidx[k] : -1 - lane);
          if (take) {
            const uint32_t minr = __reduce_min_sync(peers, rank);
            if (minr != rank) { take = false; blocked = true; bk = k; }  // look again in the next step
          }
          if (take) {
            if (t != kFree) e.steals[lane * 8 + nsteal++] = (int)t - 1;
            *tp = rank;
            const int kk = k < 4 ? k : k + 1;
            const int ax = cx + (kk % 3 - 1), ay = cy + (kk / 3 - 1);
            lbase[(g_base + n) & lmask] = (uint32_t)ax | ((uint32_t)ay << 16);
            ++n;
            const Pix pr = e.pix[nidx[k]];
            sumdx += pr.cs;
            sumdy += pr.sn;
            reg_deg = fast_atan2_deg(sumdy, sumdx);
          }
        }
        __syncwarp();
      }
      if (grow) {
        s.reg_deg = reg_deg; s.sumdx = sumdx; s.sumdy = sumdy;
        s.n = n; s.nd = nd; s.dep0 = d0; s.dep1 = d1; s.dep2 = d2; s.dep3 = d3;
        if (s.base + n > s.ext) s.ext = s.base + n;
        if (blocked) s.kst = bk;
      }
    }

    // ================= 5. steals: the victims are undone, then their dependants =================
    {
      unsigned mst = __ballot_sync(FULL, nsteal > 0);
      while (mst) {
        int owner = __ffs(mst) - 1;
        mst &= mst - 1;
        int cnt = __shfl_sync(FULL, nsteal, owner);
        for (int q = 0; q < cnt; ++q) wl_push_uniform(e, c, e.steals[owner * 8 + q]);
      }
      if (c.wn) {
        process_aborts(e, c, s);
        if (c.fail) abort_all_speculative(e, c, s);
      }
    }

    // ================= 6. FIFO advance; lanes whose list is exhausted are post-processed =================
    bool finished = false;
    if (s.st == ST_GROW && !blocked) {
      s.kst = 0;
      ++s.i;
      finished = s.i >= s.n;
    }
    unsigned mf = __ballot_sync(FULL, finished);
    while (mf) {
      int owner = __ffs(mf) - 1;
      mf &= mf - 1;
      // a release event of an earlier post-processing in this loop may have undone this one
      if (!__shfl_sync(FULL, (int)(s.st == ST_GROW), owner)) continue;
      RectCand rec;
      int done, has_cand;
      after_grow(e, c, s, owner, rec, &done, &has_cand);
      if (c.wn) {  // released pixels: dependants are re-run
        process_aborts(e, c, s);
        if (c.fail) abort_all_speculative(e, c, s);
      }
      if (!done) continue;
      int opos = __shfl_sync(FULL, s.pos, owner);
      if (opos == c.h) {
        // the head: commit at once, its entries are popped from the ring tail
        int nv = __shfl_sync(FULL, s.need_val, owner);
        bool ok = true;
        if (nv) {
          ListRef L = list_of(e, owner, __shfl_sync(FULL, s.rstart, owner), __shfl_sync(FULL, s.big, owner));
          ok = validate_walk(e, L, __shfl_sync(FULL, s.ext, owner), (uint32_t)opos + 1u);
        }
        if (!ok) {
          wl_push_uniform(e, c, opos);
          process_aborts(e, c, s);
          if (c.fail) abort_all_speculative(e, c, s);
          continue;
        }
        if (has_cand) emit_cand(e, c, cand, A.cand_cap, A.overflow, rec);
        if (lane == owner) { s.st = ST_IDLE; s.nd = 0; pop_dead_front(e, s); }
        ++c.h;
      } else {
        // park it
        if (lane == owner) {
          int sl = (s.qh + s.qn) & (kEngQ - 1);
          EngDesc* d = dslot(e, sl);
          d->pos = s.pos; d->start = s.rstart; d->ext = s.ext;
          d->flags = (has_cand ? DF_CAND : 0) | (s.need_val ? DF_VAL : 0);
          d->nd = s.nd; d->dep[0] = s.dep0; d->dep[1] = s.dep1; d->dep[2] = s.dep2; d->dep[3] = s.dep3;
          if (has_cand) e.rects[lane * kEngQ + sl] = rec;
          ++s.qn;
          s.rstart += s.ext;
          s.st = ST_IDLE; s.nd = 0;
          pop_dead_front(e, s);
        }
      }
      __syncwarp();
    }
  }
  if (lane == 0) O.n_cand[f] = c.n_cand < A.cand_cap ? c.n_cand : A.cand_cap;
}

// owner tags of a batch from its angle plane: 0 = undefined, FREE otherwise
__global__ void spec_tag_init_kernel(const float* __restrict__ ang, uint32_t* __restrict__ tag, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    tag[i] = ang[i] == kNotDefDeg ? 0u : kTagFree;
}

void launch_region_engine_spec(const EngineArgs& a, cudaStream_t st) {
  for (int o = 0; o < a.num_octaves; ++o) {
    const size_t n = (size_t)a.batch * a.oct[o].ws * a.oct[o].hs;
    spec_tag_init_kernel<<<(int)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(a.oct[o].ang, a.oct[o].tag, n);
  }
  dim3 grid(a.batch, a.num_octaves);
  region_engine_spec_kernel<<<grid, 32, 0, st>>>(a);
}

}  // namespace vpl
