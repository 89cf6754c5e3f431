#include "tmap.h"

#include <cudaTypedefs.h>
#include <stdio.h>
#include <string.h>

namespace b2n {

static thread_local char g_tmap_err[256] = "";
const char* tmap_last_error() { return g_tmap_err; }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_tiled() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}
static PFN_cuTensorMapEncodeIm2col_v12000 get_encode_im2col() {
  static PFN_cuTensorMapEncodeIm2col_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(p);
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_of(int bytes) {
  switch (bytes) {
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    case kSwizzle128Atom32: return CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}
static CUtensorMapDataType dtype_of(TmapDtype dt) {
  return dt == kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}
static uint64_t esize(TmapDtype dt) { return dt == kF16 ? 2 : 4; }

int make_tiled_map_2d(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t rows,
                      uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols,
                      int swizzle_bytes) {
  auto fn = get_encode_tiled();
  if (fn == nullptr) {
    snprintf(g_tmap_err, sizeof g_tmap_err, "cuTensorMapEncodeTiled entry point unavailable");
    return 1;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esize(dt)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype_of(dt), 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(swizzle_bytes),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tmap_err, sizeof g_tmap_err,
             "cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box=%ux%u swz=%d",
             (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld,
             box_rows, box_cols, swizzle_bytes);
    return 2;
  }
  return 0;
}

int make_grouped_map_3d(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols,
                        uint32_t box_rows, uint32_t box_groups, int swizzle_bytes) {
  auto fn = get_encode_tiled();
  if (fn == nullptr) {
    snprintf(g_tmap_err, sizeof g_tmap_err, "cuTensorMapEncodeTiled entry point unavailable");
    return 1;
  }
  cuuint64_t dims[3] = {32, rows, cols / 32};
  cuuint64_t strides[2] = {cols * 4, 128};
  cuuint32_t box[3] = {32, box_rows, box_groups};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(swizzle_bytes),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tmap_err, sizeof g_tmap_err,
             "cuTensorMapEncodeTiled(3d) failed (%d): rows=%llu cols=%llu box=%ux%u swz=%d", (int)r,
             (unsigned long long)rows, (unsigned long long)cols, box_rows, box_groups,
             swizzle_bytes);
    return 2;
  }
  return 0;
}

int make_im2col_map(CUtensorMap* out, const void* base, TmapDtype dt, int N, int H, int W, int C,
                    int R, int S, int pad_h_lo, int pad_h_hi, int pad_w_lo, int pad_w_hi,
                    int stride, uint32_t channels_per_pixel, uint32_t pixels_per_column,
                    int swizzle_bytes) {
  auto fn = get_encode_im2col();
  if (fn == nullptr) {
    snprintf(g_tmap_err, sizeof g_tmap_err, "cuTensorMapEncodeIm2col entry point unavailable");
    return 1;
  }
  const uint64_t es = esize(dt);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  int lower[2] = {-pad_w_lo, -pad_h_lo};
  int upper[2] = {pad_w_hi - (S - 1), pad_h_hi - (R - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = fn(out, dtype_of(dt), 4, const_cast<void*>(base), dims, strides, lower, upper,
                  channels_per_pixel, pixels_per_column, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_of(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tmap_err, sizeof g_tmap_err,
             "cuTensorMapEncodeIm2col failed (%d): NHWC=%d,%d,%d,%d RS=%dx%d pad=(%d,%d,%d,%d) "
             "stride=%d cpp=%u ppc=%u swz=%d es=%d",
             (int)r, N, H, W, C, R, S, pad_h_lo, pad_h_hi, pad_w_lo, pad_w_hi, stride,
             channels_per_pixel, pixels_per_column, swizzle_bytes, (int)es);
    return 2;
  }
  // Known driver issue (also worked around by CUTLASS, copy_traits_sm90_im2col.hpp): im2col
  // descriptors of tensors smaller than 128 KiB get bit 21 of word 1 set wrongly by drivers
  // <= 13.1; clear it.
  int drv = 0;
  cudaDriverGetVersion(&drv);
  const uint64_t bytes = (uint64_t)N * H * W * C * es;
  if (drv <= 13010 && bytes < 131072) {
    reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  }
  return 0;
}

}  // namespace b2n
