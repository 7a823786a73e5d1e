// lbd_match.cu -- LBD descriptor extraction and brute-force Hamming kNN (sm_100a).
//
// LBD restates BinaryDescriptor::computeLBD + binaryConversion (opencv_contrib 3.4
// modules/line_descriptor/src/binary_descriptor.cpp; CPU
// restatement oracle/orc_lbd.c lbd_one).  ONE WARP PER LINE.  The 63 rows of the
// line support region are independent until the band accumulation, so the lanes
// take rows (lane, lane+32) and each walks its row left to right with the same
// float32 running sums as the sequential code; the band sums are then accumulated
// in ascending row order.  Result: bit-identical to the sequential float32 code
// (compile with -fmad=false).  Gather-bound (L1/L2), not HBM-bound.
//
// Hamming kNN restates BinaryDescriptorMatcher::match / knnMatch as the north
// star fixes it: brute force over 256-bit codes, ascending
// distance, lowest train index on ties.  Shared-memory train tiles (word-major,
// conflict-free), XOR + POPC, per-lane top-k on a packed (distance << 20 | index)
// key, warp-shuffle min to merge.  Integer-pipe (POPC) bound.
#include <float.h>

#include "vpl_common.cuh"

namespace vpl {

constexpr int NUM_OF_BANDS = 9;
constexpr int WIDTH_OF_BAND = 7;
constexpr int LSP_H = NUM_OF_BANDS * WIDTH_OF_BAND;  // 63

__constant__ float c_gaussG[LSP_H];
__constant__ float c_gaussL[WIDTH_OF_BAND * 3];
__constant__ unsigned char c_comb[32][2];

void lbd_init_tables() {
  // gaussCoefL_/gaussCoefG_ exactly as the published code computes them: the
  // centre and sigma expressions are INTEGER divisions (u=10, sigma=7; u=sigma=31).
  float hL[WIDTH_OF_BAND * 3], hG[LSP_H];
  double u = (WIDTH_OF_BAND * 3 - 1) / 2;
  double sigma = (WIDTH_OF_BAND * 2 + 1) / 2;
  double invsigma2 = -1 / (2 * sigma * sigma);
  for (int i = 0; i < WIDTH_OF_BAND * 3; ++i) {
    double dis = i - u;
    hL[i] = (float)exp(dis * dis * invsigma2);
  }
  u = (NUM_OF_BANDS * WIDTH_OF_BAND - 1) / 2;
  sigma = u;
  invsigma2 = -1 / (2 * sigma * sigma);
  for (int i = 0; i < LSP_H; ++i) {
    double dis = i - u;
    hG[i] = (float)exp(dis * dis * invsigma2);
  }
  static const unsigned char comb[32][2] = {
      {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
      {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
      {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};
  cudaMemcpyToSymbol(c_gaussL, hL, sizeof(hL));
  cudaMemcpyToSymbol(c_gaussG, hG, sizeof(hG));
  cudaMemcpyToSymbol(c_comb, comb, sizeof(comb));
}

constexpr int LBD_WARPS = 4;

__global__ void __launch_bounds__(LBD_WARPS * 32)
lbd_kernel(LbdArgs A, const VplKeyLine* __restrict__ kl_, const int* __restrict__ counts, int cap,
           uint8_t* __restrict__ desc_, float* __restrict__ fdesc_) {
  __shared__ float s_row[LBD_WARPS][LSP_H][8];
  __shared__ float s_des[LBD_WARPS][72];
  __shared__ float s_band[LBD_WARPS][8][NUM_OF_BANDS];
  const int f = blockIdx.y;
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a frame's lines are dealt to a fixed number of warps (a grid sized by the capacity launched mostly empty CTAs)
  const int n_lines = counts[f];
  for (int li = blockIdx.x * LBD_WARPS + wi; li < n_lines; li += gridDim.x * LBD_WARPS) {
  const VplKeyLine kl = kl_[(size_t)f * cap + li];
  const int o = kl.octave;
  const int realWidth = A.w[o], realHeight = A.h[o];
  const short2* __restrict__ g = A.grad[o] + (size_t)f * realWidth * realHeight;
  const int imageWidth = realWidth - 1, imageHeight = realHeight - 1;
  const short lengthOfLSP = (short)kl.numOfPixels;
  const short halfHeight = (LSP_H - 1) / 2;
  const short halfWidth = (short)((lengthOfLSP - 1) / 2);
  const float lineMiddlePointX = 0.5f * (kl.sPointInOctaveX + kl.ePointInOctaveX);
  const float lineMiddlePointY = 0.5f * (kl.sPointInOctaveY + kl.ePointInOctaveY);
  const float dL0 = (float)cos((double)kl.angle);
  const float dL1 = (float)sin((double)kl.angle);
  const float dO0 = -dL1, dO1 = dL0;
  float sCorX0 = -dL0 * halfWidth + dL1 * halfHeight + lineMiddlePointX;
  float sCorY0 = -dL1 * halfWidth - dL0 * halfHeight + lineMiddlePointY;
  // advance the row origin to row `lane` with the same repeated float updates
  for (int t = 0; t < lane; ++t) { sCorX0 -= dL1; sCorY0 += dL0; }
  for (int hID = lane; hID < LSP_H; hID += 32) {
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdLRowSum = 0, ngdLRowSum = 0, pgdORowSum = 0, ngdORowSum = 0;
    // The sample positions do not depend on the samples: four gathers are issued before the first of them is used
    // (the row sums still take the samples one by one in wID order, the coordinates the same chain of float additions).
    auto sample = [&](void) -> short2 {
      int tx = (int)(short)roundf(sCorX);
      int xCor = (tx < 0) ? 0 : (tx > imageWidth) ? imageWidth : tx;
      int ty = (int)(short)roundf(sCorY);
      int yCor = (ty < 0) ? 0 : (ty > imageHeight) ? imageHeight : ty;
      sCorX += dL0;
      sCorY += dL1;
      return __ldg(g + yCor * realWidth + xCor);
    };
    auto accumulate = [&](short2 d) {
      float gDL = d.x * dL0 + d.y * dL1;
      float gDO = d.x * dO0 + d.y * dO1;
      if (gDL > 0) pgdLRowSum += gDL;
      else ngdLRowSum -= gDL;
      if (gDO > 0) pgdORowSum += gDO;
      else ngdORowSum -= gDO;
    };
    int wID = 0;
    for (; wID + 4 <= lengthOfLSP; wID += 4) {
      short2 d0 = sample(), d1 = sample(), d2 = sample(), d3 = sample();
      accumulate(d0); accumulate(d1); accumulate(d2); accumulate(d3);
    }
    for (; wID < lengthOfLSP; ++wID) accumulate(sample());
    float coef = c_gaussG[hID];
    pgdLRowSum = coef * pgdLRowSum;
    ngdLRowSum = coef * ngdLRowSum;
    pgdORowSum = coef * pgdORowSum;
    ngdORowSum = coef * ngdORowSum;
    float* r = s_row[wi][hID];
    r[0] = pgdLRowSum; r[1] = ngdLRowSum;
    r[2] = pgdLRowSum * pgdLRowSum; r[3] = ngdLRowSum * ngdLRowSum;
    r[4] = pgdORowSum; r[5] = ngdORowSum;
    r[6] = pgdORowSum * pgdORowSum; r[7] = ngdORowSum * ngdORowSum;
    // 32 more row steps for the second pass
    for (int t = 0; t < 32; ++t) { sCorX0 -= dL1; sCorY0 += dL0; }
  }
  __syncwarp();
  // band sums: accumulator a = q*9 + b, rows in ascending order
  const float invN2 = (float)(1.0 / (WIDTH_OF_BAND * 2.0));
  const float invN3 = (float)(1.0 / (WIDTH_OF_BAND * 3.0));
  for (int a = lane; a < 72; a += 32) {
    int q = a / NUM_OF_BANDS, b = a - q * NUM_OF_BANDS;
    bool sq = (q == 2 || q == 3 || q == 6 || q == 7);
    float acc = 0;
    int h0 = max(0, WIDTH_OF_BAND * (b - 1)), h1 = min(LSP_H, WIDTH_OF_BAND * (b + 2));
    for (int h = h0; h < h1; ++h) {
      int hb = h / WIDTH_OF_BAND, hm = h - hb * WIDTH_OF_BAND;
      int ci = (hb == b) ? hm + WIDTH_OF_BAND : (hb == b + 1) ? hm + 2 * WIDTH_OF_BAND : hm;
      float c = c_gaussL[ci];
      float v = s_row[wi][h][q];
      if (sq) acc += c * c * v;
      else acc += c * v;
    }
    s_band[wi][q][b] = acc;
  }
  __syncwarp();
  float* des = s_des[wi];
  if (lane < NUM_OF_BANDS) {
    int b = lane;
    float invN = (b == 0 || b == NUM_OF_BANDS - 1) ? invN2 : invN3;
    float temp = s_band[wi][0][b] * invN;
    des[8 * b] = temp;
    des[8 * b + 4] = sqrtf(s_band[wi][2][b] * invN - temp * temp);
    temp = s_band[wi][1][b] * invN;
    des[8 * b + 1] = temp;
    des[8 * b + 5] = sqrtf(s_band[wi][3][b] * invN - temp * temp);
    temp = s_band[wi][4][b] * invN;
    des[8 * b + 2] = temp;
    des[8 * b + 6] = sqrtf(s_band[wi][6][b] * invN - temp * temp);
    temp = s_band[wi][5][b] * invN;
    des[8 * b + 3] = temp;
    des[8 * b + 7] = sqrtf(s_band[wi][7][b] * invN - temp * temp);
  }
  __syncwarp();
  // normalisation: sequential float sums (all lanes redundantly, from shared memory)
  float tempM = 0, tempS = 0;
  for (int b = 0; b < NUM_OF_BANDS; ++b) {
    const float* d = des + 8 * b;
    tempM += d[0] * d[0]; tempM += d[1] * d[1]; tempM += d[2] * d[2]; tempM += d[3] * d[3];
    tempS += d[4] * d[4]; tempS += d[5] * d[5]; tempS += d[6] * d[6]; tempS += d[7] * d[7];
  }
  tempM = 1.0f / sqrtf(tempM);
  tempS = 1.0f / sqrtf(tempS);
  __syncwarp();
  for (int i = lane; i < 72; i += 32) {
    float v = des[i] * (((i & 7) < 4) ? tempM : tempS);
    if ((double)v > 0.4) v = (float)0.4;
    des[i] = v;
  }
  __syncwarp();
  float temp = 0;
  for (int i = 0; i < 72; ++i) temp += des[i] * des[i];
  temp = 1.0f / sqrtf(temp);
  __syncwarp();
  for (int i = lane; i < 72; i += 32) {
    des[i] = des[i] * temp;
    if (fdesc_) fdesc_[((size_t)f * cap + li) * 72 + i] = des[i];  // returnFloatDescr = true
  }
  __syncwarp();
  // binarisation: lane c -> byte c
  {
    const float* f1 = des + 8 * c_comb[lane][0];
    const float* f2 = des + 8 * c_comb[lane][1];
    unsigned r = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b)
      if (f1[b] > f2[b]) r += (1u << b);
    desc_[((size_t)f * cap + li) * 32 + lane] = (uint8_t)r;
  }
  __syncwarp();
  }  // next line of this warp
}

void launch_lbd(const LbdArgs& a, const VplKeyLine* kl, const int* counts, int cap, uint8_t* desc, float* fdesc,
                int batch, cudaStream_t st) {
  constexpr int kCtasPerFrame = 48;  // 192 lines per pass; frames with more lines take further passes
  dim3 grid(min((cap + LBD_WARPS - 1) / LBD_WARPS, kCtasPerFrame), batch);
  lbd_kernel<<<grid, LBD_WARPS * 32, 0, st>>>(a, kl, counts, cap, desc, fdesc);
}

// ---------------------------------------------------------------------------
// Hamming kNN.  CTA = 8 warps; blockIdx.y = pair, blockIdx.x = group of 32
// queries (4 per warp, processed together so each train word fetched from shared
// memory is reused 4 times).  Train codes are staged in 256-code tiles, stored
// word-major so lane t reads word w of code t without bank conflicts.
// ---------------------------------------------------------------------------
constexpr int HM_WARPS = 8, HM_QPW = 4, HM_QPB = HM_WARPS * HM_QPW, HM_TILE = 256;

template <int K>
__device__ __forceinline__ void topk_insert(unsigned (&best)[K], unsigned key) {
  if (key < best[K - 1]) {
    best[K - 1] = key;
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
      if (best[i] < best[i - 1]) {
        unsigned t = best[i]; best[i] = best[i - 1]; best[i - 1] = t;
      }
    }
  }
}

template <int K>
__global__ void __launch_bounds__(HM_WARPS * 32)
hamming_knn_kernel(const uint8_t* __restrict__ q_, const int* __restrict__ nq_, int cap_q,
                   const uint8_t* __restrict__ t_, const int* __restrict__ nt_, int cap_t, int k,
                   VplDMatch* __restrict__ out_) {
  __shared__ uint32_t s_t[8][HM_TILE + 1];
  const int pair = blockIdx.y;
  const int nq = nq_[pair], nt = nt_[pair];
  const int q0 = blockIdx.x * HM_QPB;
  if (q0 >= nq) return;
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(q_ + (size_t)pair * cap_q * 32);
  const uint32_t* t = reinterpret_cast<const uint32_t*>(t_ + (size_t)pair * cap_t * 32);
  VplDMatch* out = out_ + (size_t)pair * cap_q * k;

  uint32_t qc[HM_QPW][8];
  unsigned best[HM_QPW][K];
#pragma unroll
  for (int j = 0; j < HM_QPW; ++j) {
    int qi = q0 + wi * HM_QPW + j;
#pragma unroll
    for (int w = 0; w < 8; ++w) qc[j][w] = (qi < nq) ? __ldg(q + (size_t)qi * 8 + w) : 0u;
#pragma unroll
    for (int i = 0; i < K; ++i) best[j][i] = 0xffffffffu;
  }
  for (int t0 = 0; t0 < nt; t0 += HM_TILE) {
    __syncthreads();
    const int tn = min(HM_TILE, nt - t0);
    for (int i = threadIdx.x; i < tn * 8; i += HM_WARPS * 32) {
      int c = i >> 3, w = i & 7;
      s_t[w][c] = __ldg(t + (size_t)(t0 + c) * 8 + w);
    }
    __syncthreads();
    for (int c = lane; c < tn; c += 32) {
      uint32_t tw[8];
#pragma unroll
      for (int w = 0; w < 8; ++w) tw[w] = s_t[w][c];
#pragma unroll
      for (int j = 0; j < HM_QPW; ++j) {
        unsigned d = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) d += __popc(qc[j][w] ^ tw[w]);
        topk_insert<K>(best[j], (d << 20) | (unsigned)(t0 + c));
      }
    }
  }
  // merge the 32 per-lane lists: k rounds of warp-min; the winner pops its head
#pragma unroll
  for (int j = 0; j < HM_QPW; ++j) {
    int qi = q0 + wi * HM_QPW + j;
    for (int r = 0; r < k; ++r) {
      unsigned head = best[j][0];
      unsigned mn = head;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      if (mn != 0xffffffffu && head == mn) {
#pragma unroll
        for (int i = 0; i < K - 1; ++i) best[j][i] = best[j][i + 1];
        best[j][K - 1] = 0xffffffffu;
      }
      if (lane == 0 && qi < nq) {
        VplDMatch m;
        m.queryIdx = qi;
        m.imgIdx = 0;
        if (mn == 0xffffffffu) { m.trainIdx = -1; m.distance = FLT_MAX; }
        else { m.trainIdx = (int)(mn & 0xfffffu); m.distance = (float)(mn >> 20); }
        out[(size_t)qi * k + r] = m;
      }
    }
  }
}

// POPC throughput of the device, measured: 8 independent popc chains per thread (one POPC + one IADD per link).
__global__ void __launch_bounds__(256) popc_peak_kernel(unsigned* __restrict__ out, int iters) {
  unsigned a[8], c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * 2654435761u + k * 40503u + blockIdx.x; c[k] = a[k] >> 7; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __popc(a[k]) + c[k];
  }
  unsigned r = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) r ^= a[k];
  if (r == 0xdeadbeefu) out[0] = r;  // never true in practice: keeps the chains alive
}
void launch_popc_peak(unsigned* out, int blocks, int iters, cudaStream_t st) {
  popc_peak_kernel<<<blocks, 256, 0, st>>>(out, iters);
}

void launch_hamming_knn(const uint8_t* q, const int* nq, int cap_q, const uint8_t* t, const int* nt, int cap_t,
                        int n_pairs, int k, VplDMatch* out, cudaStream_t st) {
  dim3 grid((cap_q + HM_QPB - 1) / HM_QPB, n_pairs);
  if (k <= 1) hamming_knn_kernel<1><<<grid, HM_WARPS * 32, 0, st>>>(q, nq, cap_q, t, nt, cap_t, k, out);
  else if (k <= 2) hamming_knn_kernel<2><<<grid, HM_WARPS * 32, 0, st>>>(q, nq, cap_q, t, nt, cap_t, k, out);
  else if (k <= 4) hamming_knn_kernel<4><<<grid, HM_WARPS * 32, 0, st>>>(q, nq, cap_q, t, nt, cap_t, k, out);
  else hamming_knn_kernel<8><<<grid, HM_WARPS * 32, 0, st>>>(q, nq, cap_q, t, nt, cap_t, k, out);
}

}  // namespace vpl
