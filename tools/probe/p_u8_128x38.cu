// The CUDA programming guide's TMA example shape (libcu++ barrier + cp_async_bulk_tensor_2d_global_to_shared)
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int BW = 128, BH = 38;
__global__ void k2(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int x0, int y0) {
  __shared__ alignas(128) uint8_t s_in[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&s_in, &tmap, x0, y0, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(s_in));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = (&s_in[0][0])[i];
}
int main(int argc, char** argv) {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncFn enc = (EncFn)p;
  const int w = 752, h = 480;
  std::vector<uint8_t> img((size_t)w * h);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 7 + i / w);
  uint8_t *d, *o; cudaMalloc(&d, img.size()); cudaMalloc(&o, BW * BH); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h}; cuuint64_t gstr[1] = {(cuuint64_t)w};
  cuuint32_t box[2] = {BW, BH}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int x0 = argc > 1 ? atoi(argv[1]) : 56, y0 = argc > 2 ? atoi(argv[2]) : 29;
  k2<<<1, 128>>>(tm, o, x0, y0);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<uint8_t> got((size_t)BW * BH); cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < BH; ++r2) for (int c = 0; c < BW; ++c) {
    int gx = x0 + c, gy = y0 + r2; uint8_t want = (gx < 0 || gx >= w || gy < 0 || gy >= h) ? 0 : img[(size_t)gy * w + gx];
    bad += got[(size_t)r2 * BW + c] != want;
  }
  printf("libcu++ 2d box %dx%d at (%d,%d): encode %d launch %s mismatches %d\n", BW, BH, x0, y0, (int)r, cudaGetErrorString(e), bad);
  return 0;
}
