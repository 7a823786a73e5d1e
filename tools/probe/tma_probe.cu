// Stand-alone probe of the TMA tile load used by blur5_sobel_tma_kernel: variants of the descriptor / fences.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK, int FENCE>
__global__ void k(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int bw, int bh, int x0, int y0) {
  __shared__ __align__(128) uint8_t s_in[64 * 128];
  __shared__ __align__(8) unsigned long long s_bar;
  const uint32_t bar = smem_u32(&s_bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    if (FENCE == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bw * bh) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(s_in)),
                   "l"(&tmap), "r"(x0), "r"(y0), "r"(0), "r"(bar) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(s_in)),
                   "l"(&tmap), "r"(x0), "r"(y0), "r"(bar) : "memory");
  }
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = s_in[i];
}
int main(int argc, char** argv) {
  const int only_v = argc > 1 ? atoi(argv[1]) : -1, only_p = argc > 2 ? atoi(argv[2]) : -1;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncFn enc = (EncFn)p;
  printf("encoder %p q=%d\n", p, (int)q);
  const int w = 752, h = 480, B = 2;
  std::vector<uint8_t> img((size_t)w * h * B);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 7 + i / w);
  uint8_t *d, *o; cudaMalloc(&d, img.size()); cudaMalloc(&o, 64 * 128); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 6; ++variant) {
    if (only_v >= 0 && variant != only_v) continue;
    const int rank = (variant & 1) ? 2 : 3, fence = (variant >> 1) & 1, bw = variant >= 4 ? 128 : 80, bh = 38;
    CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B}; cuuint64_t gstr[2] = {(cuuint64_t)w, (cuuint64_t)w * h};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaMemset(o, 0xee, 64 * 128);
    for (int pos = 0; pos < 2; ++pos) {
      if (only_p >= 0 && pos != only_p) continue;
      const int x0 = pos ? 56 : -8, y0 = pos ? 29 : -3;
      if (rank == 3 && fence == 0) k<3, 0><<<1, 256>>>(tm, o, bw, bh, x0, y0);
      if (rank == 3 && fence == 1) k<3, 1><<<1, 256>>>(tm, o, bw, bh, x0, y0);
      if (rank == 2 && fence == 0) k<2, 0><<<1, 256>>>(tm, o, bw, bh, x0, y0);
      if (rank == 2 && fence == 1) k<2, 1><<<1, 256>>>(tm, o, bw, bh, x0, y0);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<uint8_t> got((size_t)bw * bh); cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r2 = 0; r2 < bh; ++r2) for (int c = 0; c < bw; ++c) {
        int gx = x0 + c, gy = y0 + r2; uint8_t want = (gx < 0 || gx >= w || gy < 0 || gy >= h) ? 0 : img[(size_t)gy * w + gx];
        bad += got[(size_t)r2 * bw + c] != want;
      }
      printf("variant %d rank %d fence %d box %dx%d pos %d: encode %d launch %s mismatches %d\n", variant, rank, fence, bw, bh, pos, (int)r, cudaGetErrorString(e), bad);
      if (e != cudaSuccess) return 0;
    }
  }
  return 0;
}
