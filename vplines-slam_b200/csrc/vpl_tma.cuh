// vpl_tma.cuh -- tile loads through the Blackwell TMA unit for the u8 stencil kernels (sm_100a).
//
// A batch of frames is described once per launch as a rank-3 tensor (x, y, frame) of bytes; a CTA then fetches its
// input tile + halo with ONE cp.async.bulk.tensor.3d that completes on an mbarrier.  Elements outside the image arrive
// as zeros; the stencils use BORDER_REFLECT_101, so border tiles overwrite them with their mirror images (which lie in
// the same tile) before the arithmetic starts.  Rules of the unit that shape the callers (measured on B200,
// tools/probe/): the box must start on a 16-byte boundary of the row (inner coordinate % 16 == 0, negative allowed),
// its width is a multiple of 16 bytes, the row pitch of the image (w) and of the frame (w*h) are multiples of 16 bytes.
// Images that do not meet this run the per-thread loads instead (same results).
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

namespace vpl {

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Thread 0 arms the barrier and issues the box copy, every thread waits for it.  `bar` needs no earlier use in the
// kernel (phase 0); dst is 128-byte aligned; bytes = box width x box height.
__device__ __forceinline__ void tma_load_box_3d(const CUtensorMap* tmap, void* dst, unsigned long long* bar_, int x, int y,
                                                int z, int bytes) {
  const uint32_t bar = smem_u32(bar_);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(0)
                 : "memory");
  }
}

// BORDER_REFLECT_101 of a tile whose element (r, c) is image pixel (gx0 + c, gy0 + r): out-of-range elements of the
// columns [c0, c0 + ncols) take the value of their mirror image, an in-range element of the same tile (sources are never
// written here: no hazard).  Called by all threads of the CTA; ends with a barrier.
__device__ __forceinline__ void tma_reflect_fix(uint8_t* tile, int pitch, int rows, int c0, int ncols, int gx0, int gy0, int w,
                                                int h, int nthreads) {
  for (int i = threadIdx.x; i < rows * ncols; i += nthreads) {
    const int r = i / ncols, c = c0 + (i - r * ncols);
    const int gx = gx0 + c, gy = gy0 + r;
    if (gx < 0 || gx >= w || gy < 0 || gy >= h) {
      const int sx = refl101(gx, w) - gx0, sy = refl101(gy, h) - gy0;
      // a mirror image outside the box belongs to a pixel no output of this tile depends on
      tile[r * pitch + c] = (sx >= 0 && sx < pitch && sy >= 0 && sy < rows) ? tile[sy * pitch + sx] : (uint8_t)0;
    }
  }
  __syncthreads();
}
#endif  // __CUDACC__

// cuTensorMapEncodeTiled through the runtime; nullptr if the driver does not have it or VPL_NO_TMA is set (measurement)
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (getenv("VPL_NO_TMA") != nullptr) return (TmapEncodeFn) nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return (TmapEncodeFn)p;
  }();
  return fn;
}

// Descriptor of `batch` u8 frames of w x h at `img` with a box of box_w x box_h x 1; false if TMA cannot describe it
inline bool make_u8_frames_tmap(CUtensorMap* tm, const uint8_t* img, int w, int h, int batch, int box_w, int box_h) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc || (w % 16) != 0 || (((size_t)w * h) % 16) != 0 || ((uintptr_t)img % 16) != 0 || w < 16 || h < 8 || (box_w % 16) != 0)
    return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
  const cuuint64_t gstr[2] = {(cuuint64_t)w, (cuuint64_t)w * h};  // bytes, dimensions 1 and 2
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(img), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace vpl
