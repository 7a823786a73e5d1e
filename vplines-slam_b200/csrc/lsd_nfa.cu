// lsd_nfa.cu -- rectangle NFA validation (rect_improve / rect_nfa / nfa) and the
// deterministic compaction + KeyLine packing (sm_100a).
//
// Restates cv::LineSegmentDetectorImpl::rect_improve / rect_nfa / nfa (opencv
// imgproc lsd.cpp, cv2 4.13 behaviour: double vertices, ceil-guarded slopes;
// SURVEY.md Appendix A.7; the nfa/log_gamma formulas are the ones the reference
// also carries at line_matching/src/edline_detector.h:210-348) and
// LSDDetector::detectImpl's KeyLine packing (opencv_contrib 3.4 LSDDetector.cpp;
// CPU restatement oracle/orc_lsd.c, oracle/orc_lbd.c).
//
// rect_improve never touches the `used` map, so every candidate rectangle of
// every frame is validated independently: ONE WARP PER CANDIDATE, lanes over the
// rows (or the columns, for flat rectangles) of the scanned rectangle.  Scans
// that share a geometry (the initial test and the five "finer precision" steps,
// and again the last five) are folded into one pass with several thresholds.
// Compaction keeps seed order: one warp per frame, ballot prefix.
#include <float.h>
#include <stdio.h>

#include "vpl_common.cuh"

namespace vpl {

__device__ __forceinline__ bool double_equal_d(double a, double b) {
  if (a == b) return true;
  double abs_diff = fabs(a - b);
  double aa = fabs(a), bb = fabs(b);
  double abs_max = aa > bb ? aa : bb;
  if (abs_max < DBL_MIN) abs_max = DBL_MIN;
  return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}

__device__ double log_gamma_lanczos_d(double x) {
  const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705,
                       1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
  double b = 0;
  for (int n = 0; n < 7; ++n) {
    a -= log(x + (double)n);
    b += q[n] * pow(x, (double)n);
  }
  return a + log(b);
}
__device__ double log_gamma_windschitl_d(double x) {
  return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
}
// (out of line on purpose: arguments beyond the table are the rare case, and inlined at the three call sites of nfa()
// these formulas were 5 400 of the kernel's 7 900 instructions, interleaved with the code that does run)
__device__ __noinline__ double log_gamma_d(double x) {
  return x > 15.0 ? log_gamma_windschitl_d(x) : log_gamma_lanczos_d(x);
}

// log_gamma of a positive integer argument: table built once on the host with the same
// Lanczos/Windschitl formulas (capi.cu); arguments beyond the table are evaluated directly.
struct LgamTab {
  const double* tab;  // tab[m] = log_gamma((double)m), m = 0..n-1 (tab[0] unused)
  int n;
};
__device__ __forceinline__ double log_gamma_int(const LgamTab& T, int m) {
  return (m < T.n) ? __ldg(T.tab + m) : log_gamma_d((double)m);
}

// The early-exit test of the binomial tail, err < tolerance * |-log10(bin_tail) - LOG_NT| * bin_tail with
// err = term * ((1 - mult^m) / (1 - mult) - 1), costs a double pow and a double log10 per iteration and decides a comparison
// that is almost never close.  This evaluates both sides in float32 -- (mult - mult^m) / (1 - mult) * (term / bin_tail)
// against tolerance * |log10(bin_tail) + LOG_NT| -- and answers only when they differ by more than 0.1 % (the float32
// evaluation is good to 1e-5, the double one to 1e-6 of the true values in the range admitted here); inside the band, or
// outside that range, the caller runs the double sequence as before.  Returns 1 (break), 0 (go on), -1 (undecided).
// Built with -DVPL_NFA_CHECK the kernel runs both and counts disagreements (vpl_debug_nfa_stats).
__device__ __forceinline__ int nfa_exit_fast(double term, double bin_tail, double mult, int m, double LOG_NT) {
  const int hi = __double2hiint(bin_tail);
  const int ex = (hi >> 20) & 0x7ff;
  if (ex == 0 || ex == 0x7ff) return -1;                       // subnormal / non-finite sum
  if (!(mult > 1e-6 && mult < 0.5) || m < 2) return -1;         // keeps the double cancellation error below 1e-6 of E
  const float mu = (float)mult;
  const float pw = exp2f((float)m * __log2f(mu));               // mult^m, relative error < 2e-3 of a term that is <= mult / 2 of E's numerator
  const float E = (mu - pw) / (1.f - mu);
  double inv;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(bin_tail));
  const float r = (float)(term * inv);                          // term / bin_tail in (0, 1]
  if (!(r > 1e-30f)) return -1;
  const unsigned mant = ((unsigned)(hi & 0xfffff) << 3) | ((unsigned)__double2loint(bin_tail) >> 29);
  const float lg10 = ((float)(ex - 1023) + __log2f(__uint_as_float(0x3f800000u | mant))) * 0.30102999566f;
  const float L = fabsf(-lg10 - (float)LOG_NT);
  if (L < 1e-2f) return -1;                                     // |log10(bin_tail) + LOG_NT| near its zero: relative error unbounded
  const float lhs = E * r, rhs = 0.1f * L;
  if (lhs < rhs * 0.999f) return 1;
  if (lhs > rhs * 1.001f) return 0;
  return -1;
}

#ifdef VPL_NFA_CHECK
__device__ unsigned long long g_nfa_stats[4];  // decisions, answered by the float32 test, disagreements, spare
#endif

// per-lane scalar; lanes may carry different (n,k,p)
__device__ __noinline__ double nfa_d(int n, int k, double p, double LOG_NT, LgamTab T) {
  if (n == 0 || k == 0) return -LOG_NT;
  if (n == k) return -LOG_NT - (double)n * log10(p);
  double p_term = p / (1 - p);
  double log1term = log_gamma_int(T, n + 1) - log_gamma_int(T, k + 1) - log_gamma_int(T, n - k + 1) +
                    (double)k * log(p) + (double)(n - k) * log(1.0 - p);
  double term = exp(log1term);
  if (double_equal_d(term, 0)) {
    if (k > n * p) return -log1term / 2.30258509299404568402 - LOG_NT;
    else return -LOG_NT;
  }
  double bin_tail = term;
  const double tolerance = 0.1;
  for (int i = k + 1; i <= n; ++i) {
    double bin_term = (double)(n - i + 1) / (double)i;
    double mult_term = bin_term * p_term;
    term *= mult_term;
    bin_tail += term;
    if (bin_term < 1) {
      const int fast = nfa_exit_fast(term, bin_tail, mult_term, n - i + 1, LOG_NT);
#ifndef VPL_NFA_CHECK
      if (fast == 1) break;
      if (fast == 0) continue;
#endif
      double err = term * ((1 - pow(mult_term, (double)(n - i + 1))) / (1 - mult_term) - 1);
      const bool stop = err < tolerance * fabs(-log10(bin_tail) - LOG_NT) * bin_tail;
#ifdef VPL_NFA_CHECK
      atomicAdd(&g_nfa_stats[0], 1ull);
      if (fast >= 0) atomicAdd(&g_nfa_stats[1], 1ull);
      if (fast >= 0 && (fast == 1) != stop) atomicAdd(&g_nfa_stats[2], 1ull);
#endif
      if (stop) break;
    }
  }
  return -log10(bin_tail) - LOG_NT;
}

struct RGeo {
  double x1, y1, x2, y2, width, dx, dy, theta;
};

// Count the pixels inside the rectangle and, for each of np thresholds, those
// aligned with theta within precs[t].  Integer results: order-free.
constexpr int NT = 6;
#ifdef VPL_DEBUG_NFA
__device__ int g_dbg_cand = -1;
__device__ int g_dbg_now = 0;
#endif
template <int NP>
__device__ __noinline__ void rect_count(const float* __restrict__ ang, int ws, int hs, const RGeo& r,
                                        const double* precs, int lane, int& total_out, int* alg_out) {
  double half_width = r.width / 2.0;
  double dyhw = r.dy * half_width;
  double dxhw = r.dx * half_width;
  double vx[4], vy[4];
  vx[0] = r.x1 - dyhw; vy[0] = r.y1 + dxhw;
  vx[1] = r.x2 - dyhw; vy[1] = r.y2 + dxhw;
  vx[2] = r.x2 + dyhw; vy[2] = r.y2 - dxhw;
  vx[3] = r.x1 + dyhw; vy[3] = r.y1 - dxhw;
  int off = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    double by = (off == 0) ? vy[0] : (off == 1) ? vy[1] : (off == 2) ? vy[2] : vy[3];
    double bx = (off == 0) ? vx[0] : (off == 1) ? vx[1] : (off == 2) ? vx[2] : vx[3];
    if (vy[i] < by || (vy[i] == by && vx[i] < bx)) off = i;
  }
  double ox[4], oy[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int s = (off + i) & 3;
    ox[i] = (s == 0) ? vx[0] : (s == 1) ? vx[1] : (s == 2) ? vx[2] : vx[3];
    oy[i] = (s == 0) ? vy[0] : (s == 1) ? vy[1] : (s == 2) ? vy[2] : vy[3];
  }
  const int c0 = (int)ceil(oy[0]), c1 = (int)ceil(oy[1]), c2 = (int)ceil(oy[2]), c3 = (int)ceil(oy[3]);
  const double flstep = (c0 != c1) ? (ox[1] - ox[0]) / (oy[1] - oy[0]) : 0;
  const double slstep = (c1 != c2) ? (ox[2] - ox[1]) / (oy[2] - oy[1]) : 0;
  const double frstep = (c0 != c3) ? (ox[3] - ox[0]) / (oy[3] - oy[0]) : 0;
  const double srstep = (c3 != c2) ? (ox[2] - ox[3]) / (oy[2] - oy[3]) : 0;
  const int ys = c0, ye = c2;
  const int n_rows = ye - ys + 1;
  int total = 0;
  int alg[NP];
#pragma unroll
  for (int t = 0; t < NP; ++t) alg[t] = 0;
  // The alignment test is decided in float32 on degrees; only inside a guard band around a
  // threshold (or the 270-degree wrap point) is the exact double sequence evaluated, so the
  // counts equal the all-double ones (float error on the difference < 1e-4 deg, band 1e-2).
  const float theta_deg = (float)(r.theta * (180.0 / VPL_PI));
  float pdeg[NP];
#pragma unroll
  for (int t = 0; t < NP; ++t) pdeg[t] = (float)(precs[t] * (180.0 / VPL_PI));
  const float GUARD = 1e-2f;

  // lanes = RL rows x XP column phases, RL the power of two (4..32) covering n_rows: tall
  // rectangles get a lane per row, flat ones several lanes along each row
  int RL = 4;
  while (RL < 32 && RL < n_rows) RL <<= 1;
  const int XP = 32 / RL;
  const int lrow = lane / XP, lph = lane - lrow * XP;
  for (int rr = lrow; rr < n_rows; rr += RL) {
    int y = ys + rr;
    if (y < 0 || y >= hs) continue;
    double left = (y <= c1) ? ox[0] + ((double)y - oy[0]) * flstep : ox[1] + ((double)y - oy[1]) * slstep;
    double right = (y < c3) ? ox[0] + ((double)y - oy[0]) * frstep : ox[3] + ((double)y - oy[3]) * srstep;
    int xs = (int)ceil(left), xe = (int)right;
    if (xs < 0) xs = 0;
    if (xe > ws - 1) xe = ws - 1;
#ifdef VPL_DEBUG_NFA
    if (g_dbg_now && lph == 0)
      printf("DBG row y=%d left=%.17g right=%.17g xs=%d xe=%d\n", y, left, right, xs, xe);
#endif
    const float* row = ang + (size_t)y * ws;
    for (int x = xs + lph; x <= xe; x += XP) {
      ++total;
      float ad = __ldg(row + x);
      if (ad >= 0.f) {
        float d = fabsf(theta_deg - ad);
        float dd = (d > 270.f) ? fabsf(d - 360.f) : d;
        bool near = fabsf(d - 270.f) <= GUARD;
#pragma unroll
        for (int t = 0; t < NP; ++t) near |= fabsf(dd - pdeg[t]) <= GUARD;
        if (!near) {
#pragma unroll
          for (int t = 0; t < NP; ++t)
            if (dd < pdeg[t]) ++alg[t];
        } else {
          double a = (double)ad * VPL_DEG2RAD;
          double n_theta = r.theta - a;
          if (n_theta < 0) n_theta = -n_theta;
          if (n_theta > VPL_3_2_PI) {
            n_theta -= VPL_2PI;
            if (n_theta < 0) n_theta = -n_theta;
          }
#pragma unroll
          for (int t = 0; t < NP; ++t)
            if (n_theta <= precs[t]) ++alg[t];
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    total += __shfl_xor_sync(0xffffffffu, total, o);
#pragma unroll
    for (int t = 0; t < NP; ++t) alg[t] += __shfl_xor_sync(0xffffffffu, alg[t], o);
  }
  total_out = total;
#pragma unroll
  for (int t = 0; t < NP; ++t) alg_out[t] = alg[t];
}

struct RState {  // the fields rect_improve mutates
  RGeo g;
  double prec, p;
};

// One "shrink" step of rect_improve's stages 2-4 (mode 0: width, 1: one side, 2: other side).
__device__ __forceinline__ bool shrink_step(RState& r, int mode) {
  const double delta = 0.5, delta_2 = delta / 2.0;
  if (!((r.g.width - delta) >= 0.5)) return false;
  if (mode == 1) {
    r.g.x1 += -r.g.dy * delta_2; r.g.y1 += r.g.dx * delta_2;
    r.g.x2 += -r.g.dy * delta_2; r.g.y2 += r.g.dx * delta_2;
  } else if (mode == 2) {
    r.g.x1 -= -r.g.dy * delta_2; r.g.y1 -= r.g.dx * delta_2;
    r.g.x2 -= -r.g.dy * delta_2; r.g.y2 -= r.g.dx * delta_2;
  }
  r.g.width -= delta;
  return true;
}

constexpr int NFA_WARPS = 4;

__global__ void __launch_bounds__(NFA_WARPS * 32)
rect_nfa_kernel(EngineArgs A) {
  const EngineOct& O = A.oct[blockIdx.z];
  const int f = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int n_cand = O.n_cand[f];
  LgamTab T;
  T.tab = A.lgam; T.n = A.lgam_n;
  const float* ang = O.ang + (size_t)f * O.ws * O.hs;
  for (int ci = blockIdx.x * NFA_WARPS + (threadIdx.x >> 5); ci < n_cand; ci += gridDim.x * NFA_WARPS) {
  RectCand* cp = O.cand + (size_t)f * A.cand_cap + ci;
  const int ws = O.ws, hs = O.hs;
  const double LOG_NT = O.log_nt;
  const double LOG_EPS = 0;
  const double delta = 0.5;

  RState rec;
  rec.g.x1 = cp->x1; rec.g.y1 = cp->y1; rec.g.x2 = cp->x2; rec.g.y2 = cp->y2;
  rec.g.width = cp->width; rec.g.dx = cp->dx; rec.g.dy = cp->dy; rec.g.theta = cp->theta;
  rec.prec = cp->prec; rec.p = cp->p;

  double log_nfa;
  {
    // initial test + "finer precision" x5 share the geometry: one scan, six thresholds
    double precs[NT], ps[NT];
    precs[0] = rec.prec; ps[0] = rec.p;
    double pp = rec.p;
#pragma unroll
    for (int t = 1; t < NT; ++t) {
      pp /= 2;
      ps[t] = pp;
      precs[t] = pp * VPL_PI;
    }
    int total, alg[NT];
#ifdef VPL_DEBUG_NFA
    if (ci == g_dbg_cand) {
      g_dbg_now = 1;
      if (lane == 0) printf("DBG cand %d geo x1=%.17g y1=%.17g x2=%.17g y2=%.17g w=%.17g dx=%.17g dy=%.17g th=%.17g\n", ci,
                            rec.g.x1, rec.g.y1, rec.g.x2, rec.g.y2, rec.g.width, rec.g.dx, rec.g.dy, rec.g.theta);
    }
    __syncwarp();
#endif
    rect_count<NT>(ang, ws, hs, rec.g, precs, lane, total, alg);
#ifdef VPL_DEBUG_NFA
    __syncwarp();
    if (ci == g_dbg_cand) {
      if (lane == 0) printf("DBG cand %d total=%d alg=%d %d %d %d %d %d\n", ci, total, alg[0], alg[1], alg[2], alg[3], alg[4], alg[5]);
      g_dbg_now = 0;
    }
    __syncwarp();
#endif
    // lane t evaluates nfa(total, alg[t], ps[t])
    int myk = 0;
    double myp = ps[0];
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (lane == t) { myk = alg[t]; myp = ps[t]; }
    double myv = nfa_d(total, myk, myp, LOG_NT, T);
    log_nfa = __shfl_sync(0xffffffffu, myv, 0);
    if (!(log_nfa > LOG_EPS)) {
      RState r = rec;
      for (int n = 0; n < 5; ++n) {
        r.p /= 2;
        r.prec = r.p * VPL_PI;
        double v = __shfl_sync(0xffffffffu, myv, n + 1);
        if (v > log_nfa) { log_nfa = v; rec = r; }
      }
    }
  }
  // Stages 2-4 (reduce width / one side / the other side).  Inside a stage the five trial
  // rectangles do not depend on the NFA outcomes (r shrinks whether or not it improved), so
  // the five scans run back to back and the five nfa() evaluations run on five lanes at once;
  // the selection then replays the sequential "if (v > log_nfa)" order.
#pragma unroll 1
  for (int mode = 0; mode < 3; ++mode) {
    if (log_nfa > LOG_EPS) break;
    RState r = rec;
    int tot[5], alg5[5];
    unsigned valid = 0;
#pragma unroll
    for (int n = 0; n < 5; ++n) {
      tot[n] = 0; alg5[n] = 0;
      if (shrink_step(r, mode)) {
        valid |= 1u << n;
        double pr[1] = {r.prec};
        int a1[1];
        rect_count<1>(ang, ws, hs, r.g, pr, lane, tot[n], a1);
        alg5[n] = a1[0];
      }
    }
    int myn = 0, myk = 0;
#pragma unroll
    for (int n = 0; n < 5; ++n)
      if (lane == n) { myn = tot[n]; myk = alg5[n]; }
    double myv = nfa_d(myn, myk, rec.p, LOG_NT, T);
    int best = -1;
#pragma unroll
    for (int n = 0; n < 5; ++n) {
      double v = __shfl_sync(0xffffffffu, myv, n);
      if (((valid >> n) & 1u) && v > log_nfa) { log_nfa = v; best = n; }
    }
    if (best >= 0) {
      // re-apply the same shrink steps: identical operations, identical rectangle
      for (int n = 0; n <= best; ++n) shrink_step(rec, mode);
    }
  }
  if (!(log_nfa > LOG_EPS)) {
    // finer precision again: geometry fixed, five thresholds in one scan
    RState r = rec;
    if ((r.g.width - delta) >= 0.5) {
      double precs[NT], ps[NT];
      double pp = r.p;
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        pp /= 2;
        ps[t] = pp;
        precs[t] = pp * VPL_PI;
      }
      ps[5] = pp;
      int total, alg[NT];
      precs[5] = -1.0;  // never counted
      rect_count<NT>(ang, ws, hs, r.g, precs, lane, total, alg);
      int myk = 0;
      double myp = ps[0];
#pragma unroll
      for (int t = 0; t < 5; ++t)
        if (lane == t) { myk = alg[t]; myp = ps[t]; }
      double myv = nfa_d(total, myk, myp, LOG_NT, T);
      for (int n = 0; n < 5; ++n) {
        r.p /= 2;
        r.prec = r.p * VPL_PI;
        double v = __shfl_sync(0xffffffffu, myv, n);
        if (v > log_nfa) { rec = r; log_nfa = v; }
      }
    }
  }
  if (lane == 0) {
    cp->x1 = rec.g.x1; cp->y1 = rec.g.y1; cp->x2 = rec.g.x2; cp->y2 = rec.g.y2;
    cp->width = rec.g.width; cp->prec = rec.prec; cp->p = rec.p;
    cp->nfa = log_nfa;
    cp->accepted = (log_nfa > LOG_EPS) ? 1 : 0;
  }
  }  // candidate loop
}

#ifdef VPL_DEBUG_NFA
void debug_set_cand(int c) { cudaMemcpyToSymbol(g_dbg_cand, &c, sizeof(int)); }
#endif

// counters of the -DVPL_NFA_CHECK build (zeros otherwise): early-exit decisions, those the float32 test answered, disagreements
void nfa_check_stats(unsigned long long out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0;
#ifdef VPL_NFA_CHECK
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_nfa_stats, 4 * sizeof(unsigned long long));
#endif
}

void launch_rect_nfa(const EngineArgs& a, cudaStream_t st) {
  // ~16 CTAs per SM in flight over (frames x octaves); warps stride over a frame's candidates
  int per_frame = (148 * 16 + a.batch * a.num_octaves - 1) / (a.batch * a.num_octaves);
  int max_pf = (a.cand_cap + NFA_WARPS - 1) / NFA_WARPS;
  if (per_frame < 1) per_frame = 1;
  if (per_frame > max_pf) per_frame = max_pf;
  dim3 grid(per_frame, a.batch, a.num_octaves);
  rect_nfa_kernel<<<grid, NFA_WARPS * 32, 0, st>>>(a);
}

// ---------------------------------------------------------------------------
// Compaction + KeyLine packing.  One warp per frame; octaves in order, candidates
// in seed order => class_id is the running index exactly as LSDDetector assigns it.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void final_segment(const RectCand& c, float e[4], double& width) {
  const double SCALE = 0.8;
  double x1 = c.x1 + 0.5, y1 = c.y1 + 0.5, x2 = c.x2 + 0.5, y2 = c.y2 + 0.5;
  x1 /= SCALE; y1 /= SCALE; x2 /= SCALE; y2 /= SCALE;
  width = c.width / SCALE;
  e[0] = (float)x1; e[1] = (float)y1; e[2] = (float)x2; e[3] = (float)y2;
}

struct PackParams {
  PackArgs a;
  float octave_scale[kMaxOctaves];
};

__global__ void __launch_bounds__(32)
pack_keylines_kernel(PackParams P, VplKeyLine* __restrict__ kl_, int* __restrict__ counts, int* overflow, int cap) {
  const int f = blockIdx.x, lane = threadIdx.x;
  const unsigned lt = (1u << lane) - 1u;
  VplKeyLine* kl = kl_ + (size_t)f * cap;
  int nout = 0;
  for (int o = 0; o < P.a.num_octaves; ++o) {
    const RectCand* cand = P.a.cand[o] + (size_t)f * P.a.cand_cap;
    const int nc = P.a.n_cand[o][f];
    const int cols = P.a.w[o], rows = P.a.h[o];
    const float octaveScale = P.octave_scale[o];
    for (int base = 0; base < nc; base += 32) {
      int j = base + lane;
      bool acc = (j < nc) && cand[j].accepted;
      unsigned m = __ballot_sync(0xffffffffu, acc);
      int pos = nout + __popc(m & lt);
      if (acc) {
        if (pos < cap) {
          float e[4];
          double width;
          final_segment(cand[j], e, width);
          // checkLineExtremes
          if (e[0] < 0) e[0] = 0;
          if (e[0] >= cols) e[0] = (float)cols - 1.0f;
          if (e[2] < 0) e[2] = 0;
          if (e[2] >= cols) e[2] = (float)cols - 1.0f;
          if (e[1] < 0) e[1] = 0;
          if (e[1] >= rows) e[1] = (float)rows - 1.0f;
          if (e[3] < 0) e[3] = 0;
          if (e[3] >= rows) e[3] = (float)rows - 1.0f;
          VplKeyLine k;
          k.startPointX = e[0] * octaveScale;
          k.startPointY = e[1] * octaveScale;
          k.endPointX = e[2] * octaveScale;
          k.endPointY = e[3] * octaveScale;
          k.sPointInOctaveX = e[0];
          k.sPointInOctaveY = e[1];
          k.ePointInOctaveX = e[2];
          k.ePointInOctaveY = e[3];
          double da = (double)(e[0] - e[2]), db = (double)(e[1] - e[3]);
          k.lineLength = (float)sqrt(da * da + db * db);
          int ix0 = __float2int_rn(e[0]), iy0 = __float2int_rn(e[1]);
          int ix1 = __float2int_rn(e[2]), iy1 = __float2int_rn(e[3]);
          int adx = abs(ix1 - ix0), ady = abs(iy1 - iy0);
          k.numOfPixels = (adx > ady ? adx : ady) + 1;
          k.angle = (float)atan2((double)(k.endPointY - k.startPointY), (double)(k.endPointX - k.startPointX));
          k.class_id = pos;
          k.octave = o;
          k.size = (k.endPointX - k.startPointX) * (k.endPointY - k.startPointY);
          k.response = k.lineLength / (float)(cols > rows ? cols : rows);
          k.pt_x = (k.endPointX + k.startPointX) / 2;
          k.pt_y = (k.endPointY + k.startPointY) / 2;
          kl[pos] = k;
        } else {
          *overflow = 1;
        }
      }
      nout += __popc(m);
    }
  }
  if (lane == 0) counts[f] = nout < cap ? nout : cap;
}

void launch_pack_keylines(const PackArgs& a, VplKeyLine* kl, int* counts, int* overflow, int cap, int batch,
                          cudaStream_t st) {
  PackParams P;
  P.a = a;
  float s = 1.0f;
  for (int o = 0; o < kMaxOctaves; ++o) {
    P.octave_scale[o] = s;  // (float)pow((float)scale, o)
    s *= (float)a.scale;
  }
  pack_keylines_kernel<<<batch, 32, 0, st>>>(P, kl, counts, overflow, cap);
}

// Raw LSD output (cv::LineSegmentDetector::detect): segments + width/prec/nfa.
__global__ void __launch_bounds__(32)
pack_segments_kernel(const RectCand* __restrict__ cand_, const int* __restrict__ n_cand, int cand_cap,
                     VplSegment* __restrict__ out_, int* __restrict__ count, int cap) {
  const int f = blockIdx.x, lane = threadIdx.x;
  const unsigned lt = (1u << lane) - 1u;
  const RectCand* cand = cand_ + (size_t)f * cand_cap;
  VplSegment* out = out_ + (size_t)f * cap;
  const int nc = n_cand[f];
  int nout = 0;
  for (int base = 0; base < nc; base += 32) {
    int j = base + lane;
    bool acc = (j < nc) && cand[j].accepted;
    unsigned m = __ballot_sync(0xffffffffu, acc);
    int pos = nout + __popc(m & lt);
    if (acc && pos < cap) {
      float e[4];
      double width;
      final_segment(cand[j], e, width);
      VplSegment s;
      s.x1 = e[0]; s.y1 = e[1]; s.x2 = e[2]; s.y2 = e[3];
      s.width = width; s.prec = cand[j].p; s.nfa = cand[j].nfa;
      out[pos] = s;
    }
    nout += __popc(m);
  }
  if (lane == 0) count[f] = nout;
}

void launch_pack_segments(const RectCand* cand, const int* n_cand, int cand_cap, VplSegment* out, int* count,
                          int cap, int batch, cudaStream_t st) {
  pack_segments_kernel<<<batch, 32, 0, st>>>(cand, n_cand, cand_cap, out, count, cap);
}

}  // namespace vpl
