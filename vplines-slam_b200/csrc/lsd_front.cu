// lsd_front.cu -- LSD stages before region growing (sm_100a):
//   K3 ll_angle : 2x2 gradient, level-line angle (cv::fastAtan2), per-frame max
//   K4 order    : pseudo-ordering = stable counting sort of the defined pixels by
//                 descending gradient bin (1024 bins), raster order inside a bin
// Restates cv::LineSegmentDetectorImpl::ll_angle (opencv imgproc lsd.cpp; SURVEY.md
// Appendix A.3; CPU restatement: oracle/orc_lsd.c ll_angle).  Exact: integer
// gradient, one float32 polynomial without FMA contraction, IEEE double sqrt/mul.
#include "vpl_common.cuh"

namespace vpl {

// gradient of the scaled image at (x,y), x<ws-1, y<hs-1
__device__ __forceinline__ void grad2x2(const uint8_t* __restrict__ s, int ws, int x, int y, int& gx, int& gy) {
  const uint8_t* r0 = s + (size_t)y * ws + x;
  const uint8_t* r1 = r0 + ws;
  int a = __ldg(r0), b = __ldg(r0 + 1), c = __ldg(r1), d = __ldg(r1 + 1);
  int DA = d - a, BC = b - c;
  gx = DA + BC;
  gy = DA - BC;
}

// ---------------------------------------------------------------------------
// K3.  Four pixels per thread.  Writes the angle (4 B) and the 16-byte engine record of every pixel and reduces
// max(gx^2+gy^2) per frame with one atomicMax per warp.  Algorithmic bytes: read S, write 8 S; the
// record is a design cost on top of that (20 S written).
// ---------------------------------------------------------------------------
constexpr int LLA_PX = 4;  // pixels per thread

__global__ void __launch_bounds__(256)
ll_angle_kernel(const uint8_t* __restrict__ scl, float* __restrict__ ang, Pix* __restrict__ pix,
                const float2* __restrict__ lut, unsigned int* __restrict__ maxq, int ws, int hs, unsigned int q_undef) {
  // pixel k of a thread is x0 + 32 k: every load/store instruction of the warp is contiguous
  const int x0 = blockIdx.x * (32 * LLA_PX) + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const size_t fo = (size_t)blockIdx.z * ws * hs;
  unsigned int q = 0;
  if (y < hs) {
    const uint8_t* r0 = scl + fo + (size_t)y * ws;
    const bool has_next_row = (y < hs - 1);
    const uint8_t* r1 = has_next_row ? r0 + ws : r0;
    int pa[LLA_PX], pb[LLA_PX], pc[LLA_PX], pd[LLA_PX];
#pragma unroll
    for (int k = 0; k < LLA_PX; ++k) {
      int xa = min(x0 + 32 * k, ws - 1), xb = min(x0 + 32 * k + 1, ws - 1);
      pa[k] = __ldg(r0 + xa); pb[k] = __ldg(r0 + xb);
      pc[k] = __ldg(r1 + xa); pd[k] = __ldg(r1 + xb);
    }
#pragma unroll
    for (int k = 0; k < LLA_PX; ++k) {
      const int x = x0 + 32 * k;
      if (x >= ws) break;
      float a = kNotDefDeg;
      uint32_t w = 0u;
      if (x < ws - 1 && has_next_row) {
        int DA = pd[k] - pa[k], BC = pb[k] - pc[k];
        w = (uint32_t)(DA + 255) | ((uint32_t)(BC + 255) << 16);
        int gx = DA + BC, gy = DA - BC;
        unsigned int qq = (unsigned int)(gx * gx + gy * gy);
        // cv2's test is norm <= rho with norm = sqrt(q / 4.0) in double.  q is an integer and the expression is monotone
        // in it, so the test is q <= q_undef with q_undef the largest q for which the double expression holds -- found
        // by evaluating that very expression on the host (launch_ll_angle): no FP64 work per pixel.
        if (qq > q_undef) {
          q = max(q, qq);
          a = fast_atan2_deg((float)gx, (float)-gy);
        }
      }
      ang[fo + (size_t)y * ws + x] = a;
      // engine record: (cos, sin) of the angle from the context's table (no FP64 work here)
      Pix u;
      u.ang = __float_as_uint(a); u.cs = 0.f; u.sn = 0.f; u.dabc = w;
      if (a != kNotDefDeg) {
        const float2 v = dabc_cssn(lut, w);
        u.cs = v.x; u.sn = v.y;
      }
      pix[fo + (size_t)y * ws + x] = u;
    }
  }
  // warp max -> one atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q = max(q, __shfl_xor_sync(0xffffffffu, q, o));
  if ((threadIdx.x & 31) == 0 && q > 0) atomicMax(maxq + blockIdx.z, q);
}

void launch_ll_angle(const uint8_t* scl, float* ang, Pix* pix, const float2* lut, unsigned int* maxq,
                     int ws, int hs, int batch, double rho, cudaStream_t st) {
  dim3 grid((ws + 32 * LLA_PX - 1) / (32 * LLA_PX), (hs + 7) / 8, batch);
  // largest integer q with sqrt((double)q / 4.0) <= rho, by the same IEEE operations the per-pixel test used (sqrt and
  // the division by 4 are correctly rounded / exact on host and device alike); q <= 2 * 510^2
  unsigned int q_undef = 0;
  while (q_undef < 520200u && sqrt((double)(int)(q_undef + 1) / 4.0) <= rho) ++q_undef;
  ll_angle_kernel<<<grid, 256, 0, st>>>(scl, ang, pix, lut, maxq, ws, hs, q_undef);
}

// (cosf, sinf) of the level-line angle as a function of the two gradient differences DA = d - a, BC = b - c of a
// pixel's 2x2 block (each in [-255, 255]): what region_grow adds to its float32 sums when it accepts the pixel,
// cos(float(angle)) / sin(float(angle)) with the float overloads = the correctly rounded float of the double
// function of the float radian angle.  One table per context (2 MB, L2-resident) instead of 8 bytes per pixel.
__global__ void cssn_lut_kernel(float2* __restrict__ lut) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kLutN * kLutN) return;
  const int DA = i / kLutN - 255, BC = i % kLutN - 255;
  const int gx = DA + BC, gy = DA - BC;
  const float a = fast_atan2_deg((float)gx, (float)-gy);
  const float ar = (float)((double)a * VPL_DEG2RAD);
  lut[i] = make_float2((float)cos((double)ar), (float)sin((double)ar));
}
void launch_cssn_lut(float2* lut, cudaStream_t st) {
  cssn_lut_kernel<<<(kLutN * kLutN + 255) / 256, 256, 0, st>>>(lut);
}

// ---------------------------------------------------------------------------
// K4.  One 512-thread CTA per frame (64 KB of counters: three CTAs share an SM; the kernel is bound by the latency of its
// row loads, so frames in flight per SM are what counts -- round 1 ran one 1024-thread CTA with 128 KB per SM).
// Warp w owns a contiguous block of rows; phase A
// computes every pixel's bin (kept as u16 in a scratch plane) and builds per-warp
// histograms in shared memory, phase B turns them into per-(warp,bin) output offsets
// (bins descending, warps ascending => raster order inside a bin), phase C re-walks
// the rows and scatters with a match_any rank.  No global atomics, deterministic.
// Algorithmic bytes: read S (u8) + 2 S write + 2 S read (u16 bins), write 4 B per defined pixel.
// ---------------------------------------------------------------------------
constexpr int ORD_THREADS = 512;
constexpr int ORD_WARPS = ORD_THREADS / 32;

__global__ void __launch_bounds__(ORD_THREADS, 3)
order_kernel(const uint8_t* __restrict__ scl, const unsigned int* __restrict__ maxq, int* __restrict__ ord,
             int* __restrict__ n_ord, uint16_t* __restrict__ bins_, size_t bins_stride, int ws, int hs, double rho) {
  extern __shared__ unsigned int s_cnt[];  // [ORD_WARPS][kBins]
  __shared__ unsigned int s_scan[ORD_WARPS];
  const int f = blockIdx.x;
  const uint8_t* s = scl + (size_t)f * ws * hs;
  int* out = ord + (size_t)f * ws * hs;
  uint16_t* bins = reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(bins_) + (size_t)f * bins_stride);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned int mq = maxq[f];
  if (mq == 0) {
    if (tid == 0) n_ord[f] = 0;
    return;
  }
  const double max_grad = sqrt((double)(int)mq / 4.0);
  const double bin_coef = (double)(kBins - 1) / max_grad;

  for (int i = tid; i < ORD_WARPS * kBins; i += ORD_THREADS) s_cnt[i] = 0;
  __syncthreads();

  // warp w owns the rows [w*R, (w+1)*R): a contiguous raster segment
  const int rows = hs - 1;
  const int R = (rows + ORD_WARPS - 1) / ORD_WARPS;
  const int y0 = wid * R, y1 = min(rows, y0 + R);
  unsigned int* mycnt = s_cnt + wid * kBins;

  // phase A: bin of every pixel (kept as u16 for phase C) + per-warp histogram
  for (int y = y0; y < y1; ++y) {
    const uint8_t* r0 = s + (size_t)y * ws;
    const uint8_t* r1 = r0 + ws;
    for (int xb = 0; xb < ws - 1; xb += 32) {
      int x = xb + lane;
      int b = -1;
      if (x < ws - 1) {
        int a = __ldg(r0 + x), bb = __ldg(r0 + x + 1), c = __ldg(r1 + x), d = __ldg(r1 + x + 1);
        int DA = d - a, BC = bb - c;
        int gx = DA + BC, gy = DA - BC;
        int qq = gx * gx + gy * gy;
        if (qq >= 100) {  // below that sqrt(q/4) <= 5 < rho: undefined without FP64 work
          double norm = sqrt((double)qq / 4.0);
          if (!(norm <= rho)) b = (int)(norm * bin_coef);
        }
        bins[(size_t)y * ws + x] = (uint16_t)b;  // 0xffff = undefined
      }
      if (b >= 0) atomicAdd(mycnt + b, 1u);  // the warp's own counters: a shared-memory atomic per defined pixel
    }
  }
  __syncthreads();

  // phase B: thread t owns the bins 2t and 2t+1.  Totals, then suffix scan over bins (descending order).
  const int b0 = 2 * tid, b1 = 2 * tid + 1;
  unsigned int t0 = 0, t1 = 0;
  for (int w = 0; w < ORD_WARPS; ++w) { t0 += s_cnt[w * kBins + b0]; t1 += s_cnt[w * kBins + b1]; }
  const unsigned int total = t0 + t1;
  unsigned int v = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned int t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  if (lane == 0) s_scan[wid] = v;  // sum of this warp's 64 bins
  __syncthreads();
  unsigned int higher = 0;  // points in the bins of higher warps
  for (int w = wid + 1; w < ORD_WARPS; ++w) higher += s_scan[w];
  const unsigned int start1 = higher + v - total;  // number of points in bins > b1
  if (tid == 0) n_ord[f] = (int)(higher + v);
  unsigned int run1 = start1, run0 = start1 + t1;   // bin b1 comes before bin b0 (descending)
  for (int w = 0; w < ORD_WARPS; ++w) {
    unsigned int c1 = s_cnt[w * kBins + b1], c0 = s_cnt[w * kBins + b0];
    s_cnt[w * kBins + b1] = run1; run1 += c1;
    s_cnt[w * kBins + b0] = run0; run0 += c0;
  }
  __syncthreads();

  // phase C: stable scatter (raster order inside a bin: rows ascend with the warp index)
  for (int y = y0; y < y1; ++y) {
    const uint16_t* br = bins + (size_t)y * ws;
    for (int xb = 0; xb < ws - 1; xb += 32) {
      int x = xb + lane;
      int b = -1;
      if (x < ws - 1) {
        unsigned int u = br[x];
        b = (u == 0xffffu) ? -1 : (int)u;
      }
      if (__ballot_sync(0xffffffffu, b >= 0) == 0) continue;
      unsigned int grp = __match_any_sync(0xffffffffu, b);
      if (b >= 0) {
        unsigned int rank = __popc(grp & ((1u << lane) - 1u));
        out[mycnt[b] + rank] = y * ws + x;
      }
      __syncwarp();
      if (b >= 0 && lane == (31 - __clz(grp))) mycnt[b] += __popc(grp);
      __syncwarp();
    }
  }
}

void launch_order(const uint8_t* scl, const unsigned int* maxq, int* ord, int* n_ord, void* scratch,
                  size_t scratch_stride, int ws, int hs, int batch, double rho, cudaStream_t st) {
  const size_t smem = (size_t)ORD_WARPS * kBins * sizeof(unsigned int);
  cudaFuncSetAttribute(order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  order_kernel<<<batch, ORD_THREADS, smem, st>>>(scl, maxq, ord, n_ord, (uint16_t*)scratch, scratch_stride, ws, hs, rho);
}

}  // namespace vpl
