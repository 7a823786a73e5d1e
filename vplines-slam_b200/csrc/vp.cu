// vp.cu -- the reference's vanishing-point stage on the device (sm_100a).
//
// Replaces vanishing_point_detection::run_vanishing_point_detection
// (/root/reference/feature_tracker/src/vanishing_point_detection.cpp:37-65, cited as vp.cpp), which
// LineFeatureTracker::readImage calls on every frame after matching
// (feature_tracker/src/line_feature_tracker.cpp:233-262).  CPU restatement: oracle/orc_vp.c
// (pinned bit for bit against the reference's own file compiled here); the device follows its
// math_mode 1 bit for bit.
//
// Stages, each one launch over all frames of the batch:
//   vp_prepare_kernel  lineinfo (vp.cpp:67-88): p1 x p2, the reference's "length"/"orientation"
//                      per line in parallel; then ONE thread replays srand(seed) and the rand()
//                      draws of getVPHypVia2Lines (:107-130) -- glibc's additive generator, whose
//                      stream is sequential -- and keeps the 105 first vanishing points vp1
//   vp_vote_kernel     getSphereGrids (:180-250), ONE WARP PER FRAME: the cells are double sums in
//                      (i, j) pair order, so the pairs are taken 32 at a time in that order, each
//                      lane computes its pair's cell and weight, and the lanes that hit the same
//                      cell (__match_any_sync) are added by their lowest lane in lane order: the
//                      sum of every cell is formed in exactly the sequential order
//   vp_vote_cta_kernel the same with ONE CTA PER FRAME for batches below 2048 frames: seven warps
//                      compute the votes of seven 32-pair chunks per round, one warp adds the
//                      previous round's chunks in order (only it touches the grid)
//   vp_smooth_kernel   the 3x3 neighbourhood pass (:252-275), one thread per cell
//   vp_score_kernel    getVPHypVia2Lines' 360 (vp2, vp3) per vp1 (:139-172) fused with
//                      getBestVpsHyp's scoring (:278-329): one thread per hypothesis, nothing
//                      is stored but the best (sum, index) per CTA; lowest index wins ties
//   vp_classify_kernel best hypothesis recomputed, vps[1]/vps[2] rule (:331-349), per-line angles
//                      in parallel, then ONE thread runs lines2Vps' sequential part (:412-497:
//                      its decisions consume the same rand() stream)
//
// Exactness.  The vote weights and the hypotheses go through atan2 / atan / sin / cos / acos; the
// oracle and the device share one deterministic definition of them (vpl_sincos.cuh, vpl_atan.cuh,
// correctly rounded in practice).  Where a transcendental only selects a grid cell, the CUDA library
// function is used first and the shared one only when the cell coordinate lies within 1e-9 of a
// cell boundary (their difference is < 1e-13), which gives the same cell for a fraction of the
// work.  The hypotheses themselves are built with the shared functions: the longitude of every vp2 is a
// whole number of degrees in exact arithmetic (see vp_cell), so its cell is decided by the last bits.
// -fmad=false, IEEE sqrt/div, sums in the reference's order.
#include <limits.h>
#include <stdlib.h>

#include "vpl_atan.cuh"
#include "vpl_common.cuh"

namespace vpl {

namespace {

constexpr double kPi = 3.1415926535897932384626433832795;  // CV_PI
constexpr int kLA = 90, kLO = 360, kCells = kLA * kLO;
constexpr int kNumVp2 = 360;
constexpr double kGuard = 1e-9;

struct GRand {  // glibc random_r TYPE_3 state (what srand()/rand() run)
  int r[31];
  int f, b;
};
__device__ __forceinline__ int grand_next(GRand& g) {
  const unsigned v = (unsigned)g.r[g.f] + (unsigned)g.r[g.b];
  g.r[g.f] = (int)v;
  if (++g.f >= 31) g.f = 0;
  if (++g.b >= 31) g.b = 0;
  return (int)(v >> 1);
}
__device__ void grand_seed(GRand& g, unsigned seed) {
  if (seed == 0) seed = 1;
  int word = (int)seed;
  g.r[0] = (int)seed;
  for (int i = 1; i < 31; ++i) {
    const long long hi = word / 127773, lo = word % 127773;
    const long long w = 16807 * lo - 2836 * hi;
    word = (int)w;
    if (word < 0) word += 2147483647;
    g.r[i] = word;
  }
  g.f = 3;
  g.b = 0;
  for (int i = 0; i < 310; ++i) (void)grand_next(g);
}

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 cross3(const V3& a, const V3& b) {
  V3 r;
  r.x = a.y * b.z - a.z * b.y;
  r.y = a.z * b.x - a.x * b.z;
  r.z = a.x * b.y - a.y * b.x;
  return r;
}
__device__ __forceinline__ void normalize_pos_z(V3& v) {  // vp.cpp:150-153 / :157-160
  if (v.z == 0.0) v.z = 0.0011;
  const double N = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  const double s = 1.0 / N;
  v.x *= s; v.y *= s; v.z *= s;
  if (v.z < 0) { v.x *= -1.0; v.y *= -1.0; v.z *= -1.0; }
}

// cell coordinate of an angle: int(angle / accuracy); `risky` when the quotient is close to an integer
// (only used where a miss falls back to the exact quotient, so the product with the reciprocal -- within 2 ulp of
// ang / acc -- serves)
__device__ __forceinline__ int cell_of(double ang, double acc, bool& risky) {
  const double q = ang * (1.0 / acc);  // acc is a compile-time constant at every call site
  const int c = (int)q;
  const double fr = q - (double)c;
  risky |= fr < kGuard || fr > 1.0 - kGuard;
  return c;
}

// ---- lineinfo + the random line pairs -----------------------------------------------------------
__global__ void __launch_bounds__(128) vp_prepare_kernel(const VplLine* __restrict__ lines, const int* __restrict__ n_lines,
                                                         int cap, VpBuffers B, VpParams P, const unsigned* __restrict__ seeds) {
  const int frame = blockIdx.x;
  const int n = n_lines[frame];
  __shared__ GRand g;
  if (threadIdx.x == 0) B.flags[frame] = 0;
  if (n < 2) {
    if (threadIdx.x == 0) B.status[frame] = -1;
    return;
  }
  const VplLine* L = lines + (size_t)frame * cap;
  double* para = B.para + (size_t)frame * cap * 3;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x1 = L[i].endpoint[0], y1 = L[i].endpoint[1], x2 = L[i].endpoint[2], y2 = L[i].endpoint[3];
    const V3 p1 = {(double)x1, (double)y1, 1.0}, p2 = {(double)x2, (double)y2, 1.0};
    const V3 c = cross3(p1, p2);
    para[3 * i] = c.x; para[3 * i + 1] = c.y; para[3 * i + 2] = c.z;
    const double dx = (double)(x1 - y1);  // float differences of these operands, as the reference writes them (:79-80)
    const double dy = (double)(x2 - y2);
    B.length[(size_t)frame * cap + i] = sqrt(dx * dx + dy * dy);
    double o = vpl_atan2_cr(dy, dx);
    if (o < 0) o += kPi;
    B.orient[(size_t)frame * cap + i] = o;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  grand_seed(g, seeds[frame]);
  long long draws = 0;
  int st = 0;
  double* vp1s = B.vp1 + (size_t)frame * P.it * 3;
  int* pairs = B.pairs + (size_t)frame * P.it * 2;
  for (int i = 0; i < P.it; ++i) {
    const int idx1 = grand_next(g) % n;
    int idx2 = grand_next(g) % n;
    draws += 2;
    while (idx2 == idx1) { idx2 = grand_next(g) % n; ++draws; }
    if (draws > P.max_draws) { st = -2; break; }
    const V3 a = {para[3 * idx1], para[3 * idx1 + 1], para[3 * idx1 + 2]};
    const V3 b = {para[3 * idx2], para[3 * idx2 + 1], para[3 * idx2 + 2]};
    const V3 v = cross3(a, b);
    if (v.z == 0) { --i; continue; }
    V3 vp1 = {v.x / v.z - P.ppx, v.y / v.z - P.ppy, P.f};
    if (vp1.z == 0) vp1.z = 0.0011;
    const double N = sqrt(vp1.x * vp1.x + vp1.y * vp1.y + vp1.z * vp1.z);
    const double s = 1.0 / N;
    vp1.x *= s; vp1.y *= s; vp1.z *= s;
    vp1s[3 * i] = vp1.x; vp1s[3 * i + 1] = vp1.y; vp1s[3 * i + 2] = vp1.z;
    pairs[2 * i] = idx1; pairs[2 * i + 1] = idx2;
  }
  int* gs = B.rng + (size_t)frame * 33;
  for (int i = 0; i < 31; ++i) gs[i] = g.r[i];
  gs[31] = g.f; gs[32] = g.b;
  B.status[frame] = st;
}

// ---- getSphereGrids: the vote -------------------------------------------------------------------
// One pair (i, j): the cell its intersection falls in (-1: no vote) and its weight (vp.cpp:212-247).
__device__ __forceinline__ void vote_pair(const V3& pi, double li, double oi, const V3& pj, double lj, double oj,
                                          const VpParams& P, int& cell, double& val) {
  const double acc = 1.0 / 180.0 * kPi;   // angelAccuracy
  const double tol = 60.0 / 180.0 * kPi;  // angelTolerance
  cell = -1;
  val = 0.0;
  const V3 pt = cross3(pi, pj);
  double dev = fabs(oi - oj);
  dev = (kPi - dev < dev) ? kPi - dev : dev;
  if (pt.z == 0 || dev > tol) return;
  const double x = pt.x / pt.z, y = pt.y / pt.z;
  const double X = x - P.ppx, Y = y - P.ppy, Z = P.f;
  const double N = sqrt(X * X + Y * Y + Z * Z);
  const double zn = Z / N;
  bool risky = false;
  double lat = acos(zn), lon = atan2(X, Y) + kPi;
  int la = cell_of(lat, acc, risky), lo = cell_of(lon, acc, risky);
  if (risky || !(lat == lat) || !(lon == lon)) {
    lat = vpl_acos_cr(zn);
    lon = vpl_atan2_cr(X, Y) + kPi;
    la = (int)(lat / acc);
    lo = (int)(lon / acc);
  }
  if (!(lat == lat) || !(lon == lon)) return;
  if (la >= kLA) la = kLA - 1;
  if (lo >= kLO) lo = kLO - 1;
  if (la < 0 || lo < 0) return;
  val = sqrt(li * lj) * (vpl_sin_cr(2.0 * dev) + 0.2);
  cell = la * kLO + lo;
}
// 32 votes of consecutive pairs added to the grid in lane (= pair) order: lanes with the same cell form a group,
// its lowest lane adds the members in lane order; groups of different cells proceed in parallel.
__device__ __forceinline__ void vote_add32(double* grid, int cell, const double* vals, int lane) {
  const unsigned valid = __ballot_sync(0xffffffffu, cell >= 0);
  if (cell >= 0) {
    const unsigned m = __match_any_sync(valid, cell);
    if ((__ffs(m) - 1) == lane) {
      double a = grid[cell];
      unsigned t = m;
      while (t) {
        const int src = __ffs(t) - 1;
        t &= t - 1;
        a += vals[src];
      }
      grid[cell] = a;
    }
  }
  __syncwarp();
}

// Variant A, ONE WARP PER FRAME (large batches: every warp computes and adds, nothing waits on a barrier).
constexpr int kVoteWarps = 4;
__global__ void __launch_bounds__(kVoteWarps * 32) vp_vote_kernel(const int* __restrict__ n_lines, int cap, VpBuffers B,
                                                                  VpParams P, int n_frames) {
  __shared__ double s_val[kVoteWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int frame = blockIdx.x * kVoteWarps + warp;
  if (frame >= n_frames || B.status[frame] != 0) return;
  const int n = n_lines[frame];
  const double* para = B.para + (size_t)frame * cap * 3;
  const double* length = B.length + (size_t)frame * cap;
  const double* orient = B.orient + (size_t)frame * cap;
  double* grid = B.grid + (size_t)frame * kCells;
  for (int i = 0; i < n - 1; ++i) {
    const V3 pi = {para[3 * i], para[3 * i + 1], para[3 * i + 2]};
    const double li = length[i], oi = orient[i];
    for (int jb = i + 1; jb < n; jb += 32) {
      const int j = jb + lane;
      int cell = -1;
      double val = 0.0;
      if (j < n) {
        const V3 pj = {para[3 * j], para[3 * j + 1], para[3 * j + 2]};
        vote_pair(pi, li, oi, pj, length[j], orient[j], P, cell, val);
      }
      s_val[warp][lane] = val;
      __syncwarp();
      vote_add32(grid, cell, s_val[warp], lane);
    }
  }
}

// Variant B, ONE CTA PER FRAME (small batches, or many lines per frame: the latency of one frame's vote is what
// counts).  Warps 1..7 compute the votes of seven consecutive 32-pair chunks per round into shared memory; warp 0
// adds the chunks of the previous round to the grid, in chunk order -- the only warp that touches the grid, so
// every cell still receives its additions in (i, j) order.  Two buffers, one barrier per round.
constexpr int kVoteCtaWarps = 8, kVoteCompute = kVoteCtaWarps - 1;
__global__ void __launch_bounds__(kVoteCtaWarps * 32) vp_vote_cta_kernel(const int* __restrict__ n_lines, int cap,
                                                                         VpBuffers B, VpParams P) {
  __shared__ double s_val[2][kVoteCompute][32];
  __shared__ int s_cell[2][kVoteCompute][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int frame = blockIdx.x;
  if (B.status[frame] != 0) return;
  const int n = n_lines[frame];
  const double* para = B.para + (size_t)frame * cap * 3;
  const double* length = B.length + (size_t)frame * cap;
  const double* orient = B.orient + (size_t)frame * cap;
  double* grid = B.grid + (size_t)frame * kCells;
  int total = 0;  // 32-pair chunks of the frame: row i has ceil((n - 1 - i) / 32)
  for (int i = 0; i < n - 1; ++i) total += (n - 1 - i + 31) >> 5;
  const int rounds = (total + kVoteCompute - 1) / kVoteCompute;
  // cursor of this compute warp: chunk number (warp - 1) of round 0
  int ci = 0, cj = 1;
  auto advance = [&](int steps) {
    for (int s_ = 0; s_ < steps && ci < n - 1; ++s_) {
      cj += 32;
      if (cj >= n) { ++ci; cj = ci + 1; }
    }
  };
  if (warp > 0) advance(warp - 1);
  for (int r = 0; r <= rounds; ++r) {
    if (warp > 0) {
      if (r < rounds) {
        int cell = -1;
        double val = 0.0;
        const int j = cj + lane;
        if (ci < n - 1 && j < n) {
          const V3 pi = {para[3 * ci], para[3 * ci + 1], para[3 * ci + 2]};
          const V3 pj = {para[3 * j], para[3 * j + 1], para[3 * j + 2]};
          vote_pair(pi, length[ci], orient[ci], pj, length[j], orient[j], P, cell, val);
        }
        s_cell[r & 1][warp - 1][lane] = cell;
        s_val[r & 1][warp - 1][lane] = val;
        advance(kVoteCompute);
      }
    } else if (r > 0) {
      const int b = (r - 1) & 1;
      for (int k = 0; k < kVoteCompute; ++k) vote_add32(grid, s_cell[b][k][lane], s_val[b][k], lane);
    }
    __syncthreads();
  }
}

// ---- 3x3 pass -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vp_smooth_kernel(VpBuffers B, int n_frames) {
  const int frame = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= kCells || B.status[frame] != 0) return;
  const double* g = B.grid + (size_t)frame * kCells;
  const int i = c / kLO, j = c - i * kLO;
  double out = 0.0;
  if (i >= 1 && i < kLA - 1 && j >= 1 && j < kLO - 1) {
    double tot = 0.0;
#pragma unroll
    for (int a = -1; a <= 1; ++a)
#pragma unroll
      for (int b = -1; b <= 1; ++b) tot += g[(i + a) * kLO + (j + b)];
    out = g[c] + tot / 9;
  }
  B.grid_new[(size_t)frame * kCells + c] = out;
}

// ---- hypotheses + scoring -----------------------------------------------------------------------
// The shared deterministic functions, out of line: the scoring kernel calls them for every hypothesis and
// stays small enough for the instruction cache (inlined everywhere it was 7000 instructions and stalled on
// instruction fetch).
__device__ __noinline__ double atan_cr_ni(double t) { return vpl_atan_cr(t); }
__device__ __noinline__ double atan2_cr_ni(double y, double x) { return vpl_atan2_cr(y, x); }
__device__ __noinline__ double acos_cr_ni(double x) { return vpl_acos_cr(x); }
__device__ __noinline__ void sincos_cr_ni(double a, double* s, double* c) { vpl_sincos_cr(a, s, c); }

// vp2 / vp3 of hypothesis (vp1, j) (vp.cpp:139-160), in the shared arithmetic.
__device__ __forceinline__ double make_hypothesis(const V3& vp1, double sl, double cl, V3& vp2, V3& vp3) {
  const double k1 = vp1.x * sl + vp1.y * cl;
  const double k2 = vp1.z;
  const double phi = atan_cr_ni(-k2 / k1);
  double Z, sp;
  sincos_cr_ni(phi, &sp, &Z);
  vp2.x = sp * sl; vp2.y = sp * cl; vp2.z = Z;
  normalize_pos_z(vp2);
  vp3 = cross3(vp1, vp2);
  normalize_pos_z(vp3);
  return phi;
}
// Sphere cell of a unit vector (vp.cpp:291-316); false = contributes nothing.  EXACT_LON: the longitude goes
// through the shared atan2 unconditionally -- vp2 = (sin phi sin lambda, sin phi cos lambda, cos phi) has
// longitude lambda_j (+ pi) = a whole number of degrees in exact arithmetic, i.e. it always sits ON a cell
// boundary and the cell the reference picks is decided by the last bits of x, y and atan2.  Otherwise the CUDA
// library function is used and the shared one only within 1e-9 of a boundary.
// lat_hint >= 0: a value known to be within 1e-12 of acos(v.z) (vp2.z = cos(phi) up to the rounding of the
// normalisation, so its latitude is |phi| unless phi is tiny) -- it replaces the library acos on the guarded path.
template <bool EXACT_LON>
__device__ __forceinline__ bool vp_cell(const V3& v, double one, int& cell, double lat_hint = -1.0) {
  if (v.z == 0.0) return false;
  bool risky = v.z > 1.0 - 1e-12;
  double lat = lat_hint >= 0.0 ? lat_hint : acos(v.z);
  int la = cell_of(lat, one, risky);
  if (risky || !(lat == lat)) {
    lat = acos_cr_ni(v.z);
    la = (int)(lat / one);
  }
  double lon;
  int lo;
  if (EXACT_LON) {
    lon = atan2_cr_ni(v.x, v.y) + kPi;
    lo = (int)(lon / one);
  } else {
    risky = false;
    lon = atan2(v.x, v.y) + kPi;
    lo = cell_of(lon, one, risky);
    if (risky || !(lon == lon)) {
      lon = atan2_cr_ni(v.x, v.y) + kPi;
      lo = (int)(lon / one);
    }
  }
  if (!(lat == lat) || !(lon == lon)) return false;
  if (la == 90) la = 89;
  if (lo == 360) lo = 359;
  if (la < 0 || la >= kLA || lo < 0 || lo >= kLO) return false;
  cell = la * kLO + lo;
  return true;
}

constexpr int kScoreThreads = 256;
// grid (splits, frames): split s scores the outer iterations [s it / splits, (s + 1) it / splits) -- a contiguous
// range of 360 (i1 - i0) hypotheses dealt to the threads round robin (360 is not a multiple of the CTA, a loop
// per outer iteration would leave half the warps waiting)
__global__ void __launch_bounds__(kScoreThreads) vp_score_kernel(VpBuffers B, VpParams P, int splits, int frame0,
                                                                 double* __restrict__ scores) {
  const int frame = frame0 + blockIdx.y, split = blockIdx.x;
  if (B.status[frame] != 0) return;
  const double* g = B.grid_new + (size_t)frame * kCells;
  const double* vp1s = B.vp1 + (size_t)frame * P.it * 3;
  const double one = 1.0 / 180.0 * kPi;
  const int i0 = P.it * split / splits, i1 = P.it * (split + 1) / splits;
  __shared__ int s_cell1[128];  // cell of vp1 per outer iteration (-1: contributes nothing); it <= 128
  for (int i = i0 + threadIdx.x; i < i1; i += kScoreThreads) {
    const V3 vp1 = {vp1s[3 * i], vp1s[3 * i + 1], vp1s[3 * i + 2]};
    int c = -1;
    if (!vp_cell<true>(vp1, one, c)) c = -1;
    s_cell1[i - i0] = c;
  }
  __syncthreads();
  double best = 0.0;
  int best_idx = INT_MAX;
  for (int idx = i0 * kNumVp2 + threadIdx.x; idx < i1 * kNumVp2; idx += kScoreThreads) {
    const int i = idx / kNumVp2, j = idx - i * kNumVp2;
    const V3 vp1 = {vp1s[3 * i], vp1s[3 * i + 1], vp1s[3 * i + 2]};
    const double sl = B.lambda_sc[2 * j], cl = B.lambda_sc[2 * j + 1];
    V3 vp2, vp3;
    const double phi = fabs(make_hypothesis(vp1, sl, cl, vp2, vp3));
    double len = 0.0;  // lineLength[i] += sphereGrid[..][..] for the three vanishing points, in this order
    int c = s_cell1[i - i0];
    if (c >= 0) len += g[c];
    if (vp_cell<true>(vp2, one, c, phi > 1e-3 ? phi : -1.0)) len += g[c];
    if (vp_cell<false>(vp3, one, c)) len += g[c];
    if (scores) scores[idx] = len;  // parity hook (vpl_debug_vp_scores): nullptr on the product path
    if (len > best) { best = len; best_idx = idx; }  // a thread's indices ascend: strict > keeps the lowest
  }
  // block arg-max, lowest index on ties
  __shared__ double s_b[kScoreThreads / 32];
  __shared__ int s_i[kScoreThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_down_sync(0xffffffffu, best, o);
    const int oi = __shfl_down_sync(0xffffffffu, best_idx, o);
    if (ob > best || (ob == best && oi < best_idx)) { best = ob; best_idx = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_b[warp] = best; s_i[warp] = best_idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kScoreThreads / 32; ++w)
      if (s_b[w] > best || (s_b[w] == best && s_i[w] < best_idx)) { best = s_b[w]; best_idx = s_i[w]; }
    B.part_best[(size_t)frame * splits + split] = best;
    B.part_idx[(size_t)frame * splits + split] = best_idx;
  }
}

// ---- best hypothesis, vps[1]/vps[2] rule, lines2Vps ---------------------------------------------
__device__ __forceinline__ float seg_angle(const VplLine& s) {  // vp.cpp:23-28: std::atan2(float, float)
  if (s.endpoint[2] > s.endpoint[0])
    return (float)vpl_atan2_cr((double)(s.endpoint[3] - s.endpoint[1]), (double)(s.endpoint[2] - s.endpoint[0]));
  return (float)vpl_atan2_cr((double)(s.endpoint[1] - s.endpoint[3]), (double)(s.endpoint[0] - s.endpoint[2]));
}

__global__ void __launch_bounds__(128) vp_classify_kernel(const VplLine* __restrict__ all_lines, const int* __restrict__ n_all,
                                                          int cap, VpBuffers B, VpParams P, int splits, int frame_count0,
                                                          double* __restrict__ vps_out, int* __restrict__ vp_idx,
                                                          double* __restrict__ line_vps) {
  const int frame = blockIdx.x;
  const int na = n_all[frame];
  const int st = B.status[frame];
  int* out_idx = vp_idx + (size_t)frame * cap;
  double* out_lv = line_vps ? line_vps + (size_t)frame * cap * 4 : nullptr;
  __shared__ double s_vps[9];
  __shared__ double s_vp2d[6];
  __shared__ GRand g;
  if (st != 0) {  // the tracker's "no vp lines" branch (line_feature_tracker.cpp:266-276): nothing is labelled
    for (int i = threadIdx.x; i < na; i += blockDim.x) {
      out_idx[i] = 3;
      if (out_lv) { out_lv[4 * i] = 0; out_lv[4 * i + 1] = 0; out_lv[4 * i + 2] = 0; out_lv[4 * i + 3] = 0; }
    }
    if (threadIdx.x < 9) vps_out[(size_t)frame * 9 + threadIdx.x] = 0.0;
    return;
  }
  if (threadIdx.x == 0) {
    double best = 0.0;
    int bi = INT_MAX;
    for (int s = 0; s < splits; ++s) {
      const double b = B.part_best[(size_t)frame * splits + s];
      const int i = B.part_idx[(size_t)frame * splits + s];
      if (b > best || (b == best && i < bi)) { best = b; bi = i; }
    }
    if (bi == INT_MAX) bi = 0;
    B.best_idx[frame] = bi;
    const int i = bi / kNumVp2, j = bi - i * kNumVp2;
    const double* v1 = B.vp1 + ((size_t)frame * P.it + i) * 3;
    const V3 vp1 = {v1[0], v1[1], v1[2]};
    V3 vp2, vp3;
    make_hypothesis(vp1, B.lambda_sc[2 * j], B.lambda_sc[2 * j + 1], vp2, vp3);
    if (frame_count0 + frame != 0 && !(fabs(vp2.y) > 0.8)) { const V3 t = vp2; vp2 = vp3; vp3 = t; }  // :337-349
    s_vps[0] = vp1.x; s_vps[1] = vp1.y; s_vps[2] = vp1.z;
    s_vps[3] = vp2.x; s_vps[4] = vp2.y; s_vps[5] = vp2.z;
    s_vps[6] = vp3.x; s_vps[7] = vp3.y; s_vps[8] = vp3.z;
    for (int k = 0; k < 3; ++k) {  // :378-385
      s_vp2d[2 * k] = s_vps[3 * k] * P.f / s_vps[3 * k + 2] + P.ppx;
      s_vp2d[2 * k + 1] = s_vps[3 * k + 1] * P.f / s_vps[3 * k + 2] + P.ppy;
    }
    const int* gs = B.rng + (size_t)frame * 33;
    for (int k = 0; k < 31; ++k) g.r[k] = gs[k];
    g.f = gs[31]; g.b = gs[32];
  }
  __syncthreads();
  if (threadIdx.x < 9) vps_out[(size_t)frame * 9 + threadIdx.x] = s_vps[threadIdx.x];
  // per line, in parallel: the three angles (:388-411) and segAngle; stored over the line-info scratch
  const VplLine* L = all_lines + (size_t)frame * cap;
  double* ang = B.para + (size_t)frame * cap * 3;
  double* sega = B.orient + (size_t)frame * cap;
  for (int i = threadIdx.x; i < na; i += blockDim.x) {
    const double x1 = L[i].endpoint[0], y1 = L[i].endpoint[1], x2 = L[i].endpoint[2], y2 = L[i].endpoint[3];
    const double xm = (x1 + x2) / 2.0, ym = (y1 + y2) / 2.0;
    double v1x = x1 - x2, v1y = y1 - y2;
    const double N1 = sqrt(v1x * v1x + v1y * v1y);
    v1x /= N1; v1y /= N1;
    for (int j = 0; j < 3; ++j) {
      double v2x = s_vp2d[2 * j] - xm, v2y = s_vp2d[2 * j + 1] - ym;
      const double N2 = sqrt(v2x * v2x + v2y * v2y);
      v2x /= N2; v2y /= N2;
      double cv = v1x * v2x + v1y * v2y;
      if (cv > 1.0) cv = 1.0;
      if (cv < -1.0) cv = -1.0;
      double a = vpl_acos_cr(cv);
      a = (kPi - a < a) ? kPi - a : a;
      ang[3 * i + j] = a;
    }
    sega[i] = (double)seg_angle(L[i]);
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  int* lx = B.lx + (size_t)frame * cap;
  int nx = 0, ny = 0, nz = 0, flags = 0;
  const double th = 1.0 / 180.0 * kPi;
  for (int i = 0; i < na; ++i) {
    double min_angle = 1000;
    int best_j = 0;
    for (int j = 0; j < 3; ++j) {
      const double a = ang[3 * i + j];
      if (a < min_angle) {
        bool flag = false;
        const int other = j == 0 ? ny : j == 1 ? nz : nx;
        if (other > 1) {
          const int idx = grand_next(g) % other;
          if (idx < nx) {
            const float cur = (float)sega[i], qry = (float)sega[lx[idx]];
            const float d = fabsf(cur - qry);
            if ((double)d < 0.175) flag = true;
          } else {
            flags |= 1;  // the reference reads lx[idx] out of range here
          }
        }
        if (!flag) {
          min_angle = a;
          best_j = j;
          if (j == 0) lx[nx++] = i;
          else if (j == 1) ++ny;
          else ++nz;
        }
      }
    }
    const int lab = min_angle < th ? best_j : 3;
    out_idx[i] = lab;
    if (out_lv) {  // line_feature_tracker.cpp:246-262
      if (lab == 3) { out_lv[4 * i] = 0; out_lv[4 * i + 1] = 0; out_lv[4 * i + 2] = 0; out_lv[4 * i + 3] = 0; }
      else {
        out_lv[4 * i] = s_vps[3 * lab]; out_lv[4 * i + 1] = s_vps[3 * lab + 1]; out_lv[4 * i + 2] = s_vps[3 * lab + 2];
        out_lv[4 * i + 3] = s_vps[3 * lab + 2] / s_vps[3 * lab + 2];
      }
    }
  }
  B.flags[frame] = flags;
}

// The PointCloud body img_callback publishes per frame (feature_tracker/src/line_feature_tracker_node.cpp:64-153):
// n x 3 point floats (undistorted first endpoint of LineFeatureTracker::undistortedLineEndPoints,
// line_feature_tracker.cpp:36-52, z = 1), then the channels id, u, v, vp_x, vp_y, vp_z, vp_z_inv (n floats
// each).  As the reference writes it, every line carries vp[i] -- the Vector4d of line number `cam` (:108-114).
__global__ void __launch_bounds__(128) vp_cloud_kernel(const VplLine* __restrict__ all_lines, const int* __restrict__ n_all,
                                                       int cap, const int* __restrict__ ids, const double* __restrict__ line_vps,
                                                       float fx, float fy, float cx, float cy, int num_of_cam, int cam,
                                                       float* __restrict__ cloud) {
  const int frame = blockIdx.x;
  const int n = min(max(n_all[frame], 0), cap);  // never past the frame's row, whatever the count buffer holds
  const VplLine* L = all_lines + (size_t)frame * cap;
  float* pts = cloud + (size_t)frame * cap * 10;
  float* ch = pts + 3 * (size_t)n;
  const double* lv = line_vps + (size_t)frame * cap * 4;
  float vp4[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) vp4[k] = cam < n ? (float)lv[4 * cam + k] : 0.0f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    pts[3 * j] = (L[j].endpoint[0] - cx) / fx;
    pts[3 * j + 1] = (L[j].endpoint[1] - cy) / fy;
    pts[3 * j + 2] = 1.0f;
    ch[j] = (float)(ids[(size_t)frame * cap + j] * num_of_cam + cam);
    ch[(size_t)n + j] = (L[j].endpoint[2] - cx) / fx;
    ch[2 * (size_t)n + j] = (L[j].endpoint[3] - cy) / fy;
#pragma unroll
    for (int k = 0; k < 4; ++k) ch[(3 + k) * (size_t)n + j] = vp4[k];
  }
}

__global__ void vp_lambda_kernel(double* sc) {  // sin / cos of j * (2 pi / 360), :99-101, :142
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= kNumVp2) return;
  const double step = 2.0 * kPi / kNumVp2;
  double s, c;
  vpl_sincos_cr(j * step, &s, &c);
  sc[2 * j] = s; sc[2 * j + 1] = c;
}

}  // namespace

void launch_vp_cloud(const VplLine* all_lines, const int* n_all, int cap, const int* ids, const double* line_vps, float fx,
                     float fy, float cx, float cy, int num_of_cam, int cam, float* cloud, int n_frames, cudaStream_t st) {
  vp_cloud_kernel<<<n_frames, 128, 0, st>>>(all_lines, n_all, cap, ids, line_vps, fx, fy, cx, cy, num_of_cam, cam, cloud);
}

void launch_vp_lambda(double* lambda_sc, cudaStream_t st) { vp_lambda_kernel<<<2, 192, 0, st>>>(lambda_sc); }

// 0: one warp per frame, 1: one CTA per frame.  VPL_VP_VOTE=0/1 forces one (measurement).
int vp_vote_variant(int n_frames) {
  static const int forced = [] { const char* e = getenv("VPL_VP_VOTE"); return e ? atoi(e) : -1; }();
  if (forced == 0 || forced == 1) return forced;
  // measured (V1, frames/s): 64 frames 85 k vs 149 k, 512 frames 213 k vs 233 k, 4096 frames 256 k vs 254 k
  return n_frames >= 2048 ? 0 : 1;
}

int vp_score_splits(int n_frames) {  // enough CTAs to fill the machine at small batches
  if (n_frames >= 1024) return 1;
  if (n_frames >= 256) return 3;
  return 15;
}

void launch_vp_prepare(const VplLine* lines, const int* n_lines, int cap, const unsigned* seeds, const VpBuffers& B,
                       const VpParams& P, int n_frames, cudaStream_t st) {
  cudaMemsetAsync(B.grid, 0, (size_t)n_frames * kCells * sizeof(double), st);
  vp_prepare_kernel<<<n_frames, 128, 0, st>>>(lines, n_lines, cap, B, P, seeds);
}
void launch_vp_vote(const int* n_lines, int cap, const VpBuffers& B, const VpParams& P, int n_frames, cudaStream_t st) {
  if (vp_vote_variant(n_frames) == 0)
    vp_vote_kernel<<<(n_frames + kVoteWarps - 1) / kVoteWarps, kVoteWarps * 32, 0, st>>>(n_lines, cap, B, P, n_frames);
  else
    vp_vote_cta_kernel<<<n_frames, kVoteCtaWarps * 32, 0, st>>>(n_lines, cap, B, P);
  vp_smooth_kernel<<<dim3((kCells + 255) / 256, n_frames), 256, 0, st>>>(B, n_frames);
}
void launch_vp_score(const VpBuffers& B, const VpParams& P, int n_frames, cudaStream_t st) {
  const int splits = vp_score_splits(n_frames);
  vp_score_kernel<<<dim3(splits, n_frames), kScoreThreads, 0, st>>>(B, P, splits, 0, nullptr);
}
// every hypothesis' score of one frame of the last batch (parity hook); leaves the per-CTA bests as they were
void launch_vp_scores_debug(const VpBuffers& B, const VpParams& P, int n_frames, int frame, double* scores, cudaStream_t st) {
  const int splits = vp_score_splits(n_frames);
  vp_score_kernel<<<dim3(splits, 1), kScoreThreads, 0, st>>>(B, P, splits, frame, scores);
}
void launch_vp_classify(const VplLine* all_lines, const int* n_all, int cap, int frame_count0, const VpBuffers& B,
                        const VpParams& P, int n_frames, double* vps, int* vp_idx, double* line_vps, cudaStream_t st) {
  vp_classify_kernel<<<n_frames, 128, 0, st>>>(all_lines, n_all, cap, B, P, vp_score_splits(n_frames), frame_count0, vps,
                                               vp_idx, line_vps);
}

}  // namespace vpl
