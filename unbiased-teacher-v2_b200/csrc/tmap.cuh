// Host-side TMA descriptor (CUtensorMap) builders. The driver entry points are resolved through
// the runtime (cudaGetDriverEntryPoint) so the library does not link libcuda directly.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace ut2 {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                     cuuint32_t, cuuint32_t, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmapApi {
  PFN_encodeTiled tiled = nullptr;
  PFN_encodeIm2col im2col = nullptr;
  int driver_version = 0;
  bool ok = false;
};

inline const TmapApi& tmap_api() {
  static TmapApi api = [] {
    TmapApi a;
    void* f1 = nullptr;
    void* f2 = nullptr;
    cudaDriverEntryPointQueryResult q1, q2;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &q1) ==
            cudaSuccess &&
        cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q2) ==
            cudaSuccess &&
        f1 && f2) {
      a.tiled = reinterpret_cast<PFN_encodeTiled>(f1);
      a.im2col = reinterpret_cast<PFN_encodeIm2col>(f2);
      cudaDriverGetVersion(&a.driver_version);
      a.ok = true;
    }
    return a;
  }();
  return api;
}

// 2-D row-major bf16 matrix [rows, cols] (cols contiguous, row stride ld elements);
// box = {box_cols (<=64 => 128 B), box_rows}, SWIZZLE_128B, OOB -> 0.
inline int make_tmap_2d_bf16(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols,
                             uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  const TmapApi& api = tmap_api();
  if (!api.ok) return -100;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = api.tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -101;
}

// im2col view of an NHWC bf16 activation [N, H, W, C] for an R x S filter with symmetric padding
// `pad` and traversal stride `stride` (dilation 1). Each load brings `pixels` consecutive output
// positions x `channels` (<= 64) input channels of one filter tap.
inline int make_tmap_im2col_bf16(CUtensorMap* m, const void* ptr, int N, int H, int W, int C,
                                 int R, int S, int stride, int pad, uint32_t channels,
                                 uint32_t pixels) {
  const TmapApi& api = tmap_api();
  if (!api.ok) return -100;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lower[2] = {-pad, -pad};                      // {W, H}
  int upper[2] = {pad - (S - 1), pad - (R - 1)};    // {W, H}
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = api.im2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims,
                          strides, lower, upper, channels, pixels, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -102;
  // Drivers up to CUDA 13.1 set a "large tensor" bit that breaks im2col maps of tensors smaller
  // than 128 KiB; clear it (same workaround as the CUTLASS im2col descriptor path).
  if (api.driver_version <= 13010) {
    uint64_t bytes = (uint64_t)N * H * W * C * 2;
    if (bytes < 131072) m->opaque[1] &= ~(1ull << 21);
  }
  return 0;
}

}  // namespace ut2
