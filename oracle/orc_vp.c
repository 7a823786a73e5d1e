/* CPU oracle (TEST INFRASTRUCTURE ONLY -- see vpl_oracle.h): restatement of the reference's
 * vanishing-point stage, vanishing_point_detection::run_vanishing_point_detection
 * (feature_tracker/src/vanishing_point_detection.cpp:37-65), SURVEY.md 8f-4:
 *
 *   lineinfo            :67-88    line parameters p1 x p2, "length" and "orientation" (with the
 *                                 reference's own operands: dx = x1 - y1, dy = x2 - y2)
 *   getVPHypVia2Lines   :89-177   105 x 360 hypotheses from random line pairs (glibc rand())
 *   getSphereGrids      :180-276  pair intersections voted into a 90 x 360 grid + 3x3 smoothing
 *   getBestVpsHyp       :278-366  best hypothesis by the sum of its three cells, vps[1]/vps[2] order
 *   lines2Vps           :368-500  line classification (draws from the same rand() stream)
 *
 * The reference seeds the generator with time(NULL) (:107); here the seed is an argument (what
 * time() returned) and rand()/srand() are glibc's TYPE_3 additive generator restated (orc_grand_*),
 * checked against the C library's in tests/test_oracle_vp.py.
 *
 * Two repairs, both flagged so that a comparison can skip such frames:
 *   - lines2Vps reads lx[idx] with idx drawn from ly.size() / lz.size() (:438-441, :456-459); where
 *     idx >= lx.size() the reference reads outside the vector.  Here the draw is consumed, no query
 *     is made (flag stays false) and ORC_VP_FLAG_LX_OOB is reported.
 *   - a line pair whose intersection has z == 0 is redrawn for ever if no other pair exists (:126-130);
 *     after ORC_VP_MAX_DRAWS draws the function gives up with -2.
 *
 * math_mode 0 uses libm (bit-equal to the reference's own code compiled here, oracle/_ref);
 * math_mode 1 uses the shared deterministic functions of orc_sincos.h / orc_atan.h, the definition
 * the device follows bit for bit (tools/gen_atan.py says why). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "orc_atan.h"
#include "vpl_oracle.h"

double orc_atan2_cr(double y, double x) { return vpl_atan2_cr(y, x); }
double orc_atan_cr(double t) { return vpl_atan_cr(t); }
double orc_acos_cr(double x) { return vpl_acos_cr(x); }
double orc_sin_cr(double a) { return vpl_sin_cr(a); }
void orc_sincos_cr(double a, double* s, double* c) { vpl_sincos_cr(a, s, c); }

/* ---- glibc srandom_r / random_r, TYPE_3 (degree 31, separation 3): what srand()/rand() run ---- */
void orc_grand_seed(OrcGRand* g, unsigned seed) {
  if (seed == 0) seed = 1;
  int32_t word = (int32_t)seed; /* glibc keeps the running word in an int32_t: seeds >= 2^31 go negative */
  g->r[0] = (int32_t)seed;
  for (int i = 1; i < 31; ++i) {
    int64_t hi = word / 127773, lo = word % 127773;
    int64_t w = 16807 * lo - 2836 * hi;
    word = (int32_t)w;
    if (word < 0) word += 2147483647;
    g->r[i] = word;
  }
  g->f = 3;
  g->b = 0;
  for (int i = 0; i < 310; ++i) (void)orc_grand_next(g);
}
int32_t orc_grand_next(OrcGRand* g) {
  uint32_t v = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
  g->r[g->f] = (int32_t)v;
  if (++g->f >= 31) g->f = 0;
  if (++g->b >= 31) g->b = 0;
  return (int32_t)(v >> 1);
}

typedef struct { double x, y, z; } V3;
static V3 cross3(V3 a, V3 b) { /* Eigen's cross product: plain products and differences */
  V3 r;
  r.x = a.y * b.z - a.z * b.y;
  r.y = a.z * b.x - a.x * b.z;
  r.z = a.x * b.y - a.y * b.x;
  return r;
}

typedef struct { int mode; } M;
static double m_sin(const M* m, double a) { if (!m->mode) return sin(a); double s, c; vpl_sincos_cr(a, &s, &c); return s; }
static double m_cos(const M* m, double a) { if (!m->mode) return cos(a); double s, c; vpl_sincos_cr(a, &s, &c); return c; }
static double m_atan(const M* m, double a) { return m->mode ? vpl_atan_cr(a) : atan(a); }
static double m_atan2(const M* m, double y, double x) { return m->mode ? vpl_atan2_cr(y, x) : atan2(y, x); }
static double m_acos(const M* m, double a) { return m->mode ? vpl_acos_cr(a) : acos(a); }
/* std::atan2(float, float) is the float overload (segAngle, :23-28) */
static float m_atan2f(const M* m, float y, float x) { return m->mode ? (float)vpl_atan2_cr((double)y, (double)x) : atan2f(y, x); }

#define PI_CV 3.1415926535897932384626433832795 /* CV_PI */

static float seg_angle(const M* m, const OrcLine* s) { /* :23-28, returned double, stored in a float by the caller */
  if (s->endpoint[2] > s->endpoint[0]) return m_atan2f(m, s->endpoint[3] - s->endpoint[1], s->endpoint[2] - s->endpoint[0]);
  return m_atan2f(m, s->endpoint[1] - s->endpoint[3], s->endpoint[0] - s->endpoint[2]);
}

int orc_vp_hypothesis_count(void) { /* :93-97 */
  double noiseRatio = 0.5;
  double p = 1.0 / 3.0 * pow(1.0 - noiseRatio, 2);
  double confEfficience = 0.9999;
  int it = (int)(log(1 - confEfficience) / log(1.0 - p));
  return it;
}

/* One hypothesis (vp1 from the pair, vp2/vp3 from the angle index j), :118-172.  out: 9 doubles. */
static void vp_hypothesis(const M* m, V3 vp1, int j, double* out) {
  const int numVp2 = 360;
  const double stepVp2 = 2.0 * PI_CV / numVp2;
  double lambda = j * stepVp2;
  double sl = m_sin(m, lambda), cl = m_cos(m, lambda);
  double k1 = vp1.x * sl + vp1.y * cl;
  double k2 = vp1.z;
  double phi = m_atan(m, -k2 / k1);
  double Z = m_cos(m, phi);
  double sp = m_sin(m, phi);
  double X = sp * sl;
  double Y = sp * cl;
  V3 vp2 = {X, Y, Z};
  if (vp2.z == 0.0) vp2.z = 0.0011;
  double N = sqrt(vp2.x * vp2.x + vp2.y * vp2.y + vp2.z * vp2.z);
  double s = 1.0 / N;
  vp2.x *= s; vp2.y *= s; vp2.z *= s;
  if (vp2.z < 0) { vp2.x *= -1.0; vp2.y *= -1.0; vp2.z *= -1.0; }
  V3 vp3 = cross3(vp1, vp2);
  if (vp3.z == 0.0) vp3.z = 0.0011;
  N = sqrt(vp3.x * vp3.x + vp3.y * vp3.y + vp3.z * vp3.z);
  s = 1.0 / N;
  vp3.x *= s; vp3.y *= s; vp3.z *= s;
  if (vp3.z < 0) { vp3.x *= -1.0; vp3.y *= -1.0; vp3.z *= -1.0; }
  out[0] = vp1.x; out[1] = vp1.y; out[2] = vp1.z;
  out[3] = vp2.x; out[4] = vp2.y; out[5] = vp2.z;
  out[6] = vp3.x; out[7] = vp3.y; out[8] = vp3.z;
}

int orc_vp_detect(const OrcLine* lines, int n_lines, const OrcLine* all_lines, int n_all, float f_, float cx_,
                  float cy_, unsigned seed, int frame_count, int math_mode, double* vps, int32_t* vp_idx,
                  double* grid_out, int32_t* best_idx, int32_t* pairs_out, int32_t* flags, double* scores_out) {
  const M mm = {math_mode}; const M* m = &mm;
  const double f = f_, ppx = cx_, ppy = cy_; /* init(): float arguments stored in doubles (:29-34) */
  if (flags) *flags = 0;
  if (n_lines < 2 || n_all < 0) return -1;
  const int num = n_lines;
  V3* para = (V3*)malloc(sizeof(V3) * (size_t)num);
  double* length = (double*)malloc(sizeof(double) * (size_t)num);
  double* orient = (double*)malloc(sizeof(double) * (size_t)num);
  /* lineinfo :67-88 */
  for (int i = 0; i < num; ++i) {
    V3 p1 = {lines[i].endpoint[0], lines[i].endpoint[1], 1.0};
    V3 p2 = {lines[i].endpoint[2], lines[i].endpoint[3], 1.0};
    para[i] = cross3(p1, p2);
    double dx = lines[i].endpoint[0] - lines[i].endpoint[1]; /* float subtraction, as written */
    double dy = lines[i].endpoint[2] - lines[i].endpoint[3];
    length[i] = sqrt(dx * dx + dy * dy);
    double o = m_atan2(m, dy, dx);
    if (o < 0) o += PI_CV;
    orient[i] = o;
  }

  /* getVPHypVia2Lines :89-177: the pairs first (sequential draws), the hypotheses are scored below */
  const int it = orc_vp_hypothesis_count();
  const int numVp2 = 360;
  V3* vp1s = (V3*)malloc(sizeof(V3) * (size_t)it);
  OrcGRand g;
  orc_grand_seed(&g, seed);
  int64_t draws = 0;
  int rc = 0;
  for (int i = 0; i < it; ++i) {
    int idx1 = orc_grand_next(&g) % num;
    int idx2 = orc_grand_next(&g) % num;
    draws += 2;
    while (idx2 == idx1) { idx2 = orc_grand_next(&g) % num; ++draws; }
    if (draws > ORC_VP_MAX_DRAWS) { rc = -2; break; }
    V3 v = cross3(para[idx1], para[idx2]);
    if (v.z == 0) { --i; continue; }
    V3 vp1 = {v.x / v.z - ppx, v.y / v.z - ppy, f};
    if (vp1.z == 0) vp1.z = 0.0011;
    double N = sqrt(vp1.x * vp1.x + vp1.y * vp1.y + vp1.z * vp1.z);
    double s = 1.0 / N;
    vp1.x *= s; vp1.y *= s; vp1.z *= s;
    vp1s[i] = vp1;
    if (pairs_out) { pairs_out[2 * i] = idx1; pairs_out[2 * i + 1] = idx2; }
  }
  if (rc) { free(para); free(length); free(orient); free(vp1s); return rc; }

  /* getSphereGrids :180-276 */
  const double angelAccuracy = 1.0 / 180.0 * PI_CV;
  const int gridLA = (int)((PI_CV / 2.0) / angelAccuracy);
  const int gridLO = (int)((PI_CV * 2.0) / angelAccuracy);
  double* grid = (double*)calloc((size_t)gridLA * gridLO, sizeof(double));
  double* gridNew = (double*)calloc((size_t)gridLA * gridLO, sizeof(double));
  const double angelTolerance = 60.0 / 180.0 * PI_CV;
  for (int i = 0; i < num - 1; ++i) {
    for (int j = i + 1; j < num; ++j) {
      V3 pt = cross3(para[i], para[j]);
      if (pt.z == 0) continue;
      double x = pt.x / pt.z, y = pt.y / pt.z;
      double X = x - ppx, Y = y - ppy, Z = f;
      double N = sqrt(X * X + Y * Y + Z * Z);
      double latitude = m_acos(m, Z / N);
      double longitude = m_atan2(m, X, Y) + PI_CV;
      int LA = (int)(latitude / angelAccuracy);
      if (LA >= gridLA) LA = gridLA - 1;
      int LO = (int)(longitude / angelAccuracy);
      if (LO >= gridLO) LO = gridLO - 1;
      double angleDev = fabs(orient[i] - orient[j]);
      angleDev = (PI_CV - angleDev < angleDev) ? PI_CV - angleDev : angleDev; /* std::min(a, b): b < a ? b : a */
      if (angleDev > angelTolerance) continue;
      if (LA < 0 || LO < 0) continue; /* NaN intersection (overflowing products): the reference would index out of range */
      grid[LA * gridLO + LO] += sqrt(length[i] * length[j]) * (m_sin(m, 2.0 * angleDev) + 0.2);
    }
  }
  const int halfSize = 1, winSize = 3, neighNum = 9;
  for (int i = halfSize; i < gridLA - halfSize; ++i)
    for (int j = halfSize; j < gridLO - halfSize; ++j) {
      double neighborTotal = 0.0;
      for (int a = 0; a < winSize; ++a)
        for (int b = 0; b < winSize; ++b) neighborTotal += grid[(i - halfSize + a) * gridLO + (j - halfSize + b)];
      gridNew[i * gridLO + j] = grid[i * gridLO + j] + neighborTotal / neighNum;
    }
  if (grid_out) memcpy(grid_out, gridNew, sizeof(double) * (size_t)gridLA * gridLO);

  /* getBestVpsHyp :278-366 */
  const double oneDegree = 1.0 / 180.0 * PI_CV;
  int bestIdx = 0;
  double maxLength = 0.0;
  double hyp[9], best[9];
  memset(best, 0, sizeof(best));
  for (int i = 0; i < it * numVp2; ++i) {
    vp_hypothesis(m, vp1s[i / numVp2], i % numVp2, hyp);
    double lineLength = 0.0;
    for (int j = 0; j < 3; ++j) {
      const double* v = hyp + 3 * j;
      if (v[2] == 0.0) continue;
      double latitude = m_acos(m, v[2]);
      double longitude = m_atan2(m, v[0], v[1]) + PI_CV;
      if (!(latitude == latitude) || !(longitude == longitude)) continue; /* |z| > 1 by an ulp: see header */
      int la = (int)(latitude / oneDegree);
      if (la == 90) la = 89;
      int lo = (int)(longitude / oneDegree);
      if (lo == 360) lo = 359;
      if (la < 0 || la >= gridLA || lo < 0 || lo >= gridLO) continue; /* z < 0 (only with f < 0): out of the grid */
      lineLength += gridNew[la * gridLO + lo];
    }
    if (scores_out) scores_out[i] = lineLength;
    if (i == 0) memcpy(best, hyp, sizeof(best));
    if (lineLength > maxLength) { maxLength = lineLength; bestIdx = i; memcpy(best, hyp, sizeof(best)); }
  }
  if (best_idx) *best_idx = bestIdx;
  /* :331-349: row_f is 1 whenever it is read; the function-static first result is never used */
  if (frame_count != 0) {
    int row_v = fabs(best[4]) > 0.8 ? 1 : 2;
    if (row_v != 1)
      for (int k = 0; k < 3; ++k) { double t = best[3 + k]; best[3 + k] = best[6 + k]; best[6 + k] = t; }
  }
  memcpy(vps, best, sizeof(best));

  /* lines2Vps :368-500 */
  const double thAngle = 1.0 / 180.0 * PI_CV;
  double vp2Dx[3], vp2Dy[3];
  for (int i = 0; i < 3; ++i) {
    vp2Dx[i] = best[3 * i] * f / best[3 * i + 2] + ppx;
    vp2Dy[i] = best[3 * i + 1] * f / best[3 * i + 2] + ppy;
  }
  int* lx = (int*)malloc(sizeof(int) * (size_t)(3 * n_all + 3));
  int* ly = lx + n_all + 1;
  int* lz = ly + n_all + 1;
  int nx = 0, ny = 0, nz = 0;
  for (int i = 0; i < n_all; ++i) {
    double x1 = all_lines[i].endpoint[0], y1 = all_lines[i].endpoint[1];
    double x2 = all_lines[i].endpoint[2], y2 = all_lines[i].endpoint[3];
    double xm = (x1 + x2) / 2.0, ym = (y1 + y2) / 2.0;
    double v1x = x1 - x2, v1y = y1 - y2;
    double N1 = sqrt(v1x * v1x + v1y * v1y);
    v1x /= N1; v1y /= N1;
    double minAngle = 1000;
    int bestJ = 0;
    for (int j = 0; j < 3; ++j) {
      double v2x = vp2Dx[j] - xm, v2y = vp2Dy[j] - ym;
      double N2 = sqrt(v2x * v2x + v2y * v2y);
      v2x /= N2; v2y /= N2;
      double crossValue = v1x * v2x + v1y * v2y;
      if (crossValue > 1.0) crossValue = 1.0;
      if (crossValue < -1.0) crossValue = -1.0;
      double angle = m_acos(m, crossValue);
      angle = (PI_CV - angle < angle) ? PI_CV - angle : angle;
      if (angle < minAngle) {
        int flag = 0;
        const int other = j == 0 ? ny : j == 1 ? nz : nx; /* the list whose size gates and bounds the draw */
        if (other > 1) {
          int idx = orc_grand_next(&g) % other;
          if (idx < nx) {
            float cur_angle = seg_angle(m, &all_lines[i]);
            float query_angle = seg_angle(m, &all_lines[lx[idx]]);
            float delta_angle = fabsf(cur_angle - query_angle);
            if (delta_angle < 0.175) flag = 1;
          } else if (flags) {
            *flags |= ORC_VP_FLAG_LX_OOB;
          }
        }
        if (!flag) {
          minAngle = angle;
          bestJ = j;
          if (j == 0) lx[nx++] = i;
          else if (j == 1) ly[ny++] = i;
          else lz[nz++] = i;
        }
      }
    }
    vp_idx[i] = minAngle < thAngle ? bestJ : 3;
  }
  free(lx); free(grid); free(gridNew); free(para); free(length); free(orient); free(vp1s);
  return 0;
}

/* n_frames frames over n_threads threads is done by the caller (bench.py); this runs a chunk. */
int64_t orc_vp_sequence(const OrcLine* lines, const int32_t* counts, int n_frames, int cap, float f, float cx, float cy,
                        const uint32_t* seeds, int frame_count0, int math_mode, double* vps, int32_t* vp_idx) {
  int64_t labelled = 0;
  for (int i = 0; i < n_frames; ++i) {
    int32_t fl;
    if (counts[i] < 3) continue;
    if (orc_vp_detect(lines + (size_t)i * cap, counts[i], lines + (size_t)i * cap, counts[i], f, cx, cy, seeds[i],
                      frame_count0 + i, math_mode, vps + 9 * (size_t)i, vp_idx + (size_t)i * cap, 0, 0, 0, &fl, 0) == 0)
      for (int k = 0; k < counts[i]; ++k) labelled += vp_idx[(size_t)i * cap + k] != 3;
  }
  return labelled;
}

/* The sensor_msgs::PointCloud img_callback publishes per frame (feature_tracker/src/line_feature_tracker_node.cpp:
 * 64-153), for camera `cam` (the loop index i of :88), as the message body lays the numbers out: points (n x 3
 * floats: undistorted first endpoint, z = 1), then the seven channels id_of_line, u_of_endpoint, v_of_endpoint,
 * vp_x, vp_y, vp_z, vp_z_inv (n floats each).  The endpoints are LineFeatureTracker::undistortedLineEndPoints
 * (line_feature_tracker.cpp:36-52: (x - cx) / fx in float).  As written in the reference, every line carries
 * vp[i] -- the Vector4d of line number `cam`, not of line j (:108-114); line_vps == NULL or n_vps == 0 is the
 * `vp.empty()` branch.  PARITY UNPINNED for the layout (ROS message types are not in this image): this restates
 * the loop. */
void orc_line_cloud(const OrcLine* lines, const int32_t* ids, int n, const double* line_vps, int n_vps, float fx,
                    float fy, float cx, float cy, int num_of_cam, int cam, float* cloud) {
  float* pts = cloud;
  float* ch = cloud + 3 * (size_t)n;
  for (int j = 0; j < n; ++j) {
    pts[3 * j] = (lines[j].endpoint[0] - cx) / fx;
    pts[3 * j + 1] = (lines[j].endpoint[1] - cy) / fy;
    pts[3 * j + 2] = 1;
    ch[0 * (size_t)n + j] = (float)(ids[j] * num_of_cam + cam);
    ch[1 * (size_t)n + j] = (lines[j].endpoint[2] - cx) / fx;
    ch[2 * (size_t)n + j] = (lines[j].endpoint[3] - cy) / fy;
    for (int k = 0; k < 4; ++k)
      ch[(3 + k) * (size_t)n + j] = (line_vps && n_vps > 0 && cam < n_vps) ? (float)line_vps[4 * cam + k] : 0.0f;
  }
}
