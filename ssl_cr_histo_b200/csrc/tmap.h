// Host-side TMA tensor-map construction.  The driver entry points are resolved at run time
// through cudaGetDriverEntryPoint, so libb2n.so has no link-time dependency on libcuda and
// loads (for symbol checks) on a machine without a GPU driver.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2n {

// swizzle selector values: 0 (none), 32, 64, 128 (16-byte chunks) or kSwizzle128Atom32
// (128-byte span swizzled in 32-byte chunks -- required for MN-major TF32 operands).
constexpr int kSwizzle128Atom32 = 129;

// element types of the mapped tensors
enum TmapDtype { kF32 = 0, kF16 = 1 };

// 2-D row-major matrix [rows][cols] (cols contiguous, row pitch = ld elements);
// box = box_cols x box_rows.  box_cols * element size must equal the swizzle span.
int make_tiled_map_2d(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t rows,
                      uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols,
                      int swizzle_bytes);

// Row-major fp32 matrix [rows][cols] viewed as (32-column group, row, group index) so that one
// box fetches `box_groups` consecutive 32-column groups of `box_rows` rows, laid out in shared
// memory group-major ([group][row][32 cols]) -- the MN-major operand layout of the wgrad kernel.
int make_grouped_map_3d(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols,
                        uint32_t box_rows, uint32_t box_groups, int swizzle_bytes);

// NHWC activation tensor viewed in im2col mode.  Bounding box of base pixels is
// [-pad_lo, dim - 1 + pad_hi - (taps - 1)] per spatial axis; traversal stride = conv stride.
int make_im2col_map(CUtensorMap* out, const void* base, TmapDtype dt, int N, int H, int W, int C,
                    int R, int S, int pad_h_lo, int pad_h_hi, int pad_w_lo, int pad_w_hi,
                    int stride, uint32_t channels_per_pixel, uint32_t pixels_per_column,
                    int swizzle_bytes);

const char* tmap_last_error();

}  // namespace b2n
