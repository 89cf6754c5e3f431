// Weight-gradient implicit GEMM on tcgen05 (sm_100a), TF32 operands / FP32 accumulate.
//
//   dW[k][tap*Cin + c] = sum_{pixels m} X[pixel(m) + tap][c] * dY[m][k]
//
// GEMM view: D[row = (tap,c)][col = k], reduction over output pixels.  Both operands are
// "MN-major" for the tensor core (the reduction index -- the pixel -- is the slow smem axis).
// MN-major TF32 operands must use the 128B-span / 32B-atom swizzle (UMMA layout type
// SWIZZLE_128B_BASE32B, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms are 32 channels x 4
// pixels (512 B), one K=8 MMA consumes two of them.
//   * A rows: im2col-mode TMA boxes of (32 channels x PX pixels), one per 32-channel group;
//   * B cols: tiled-mode TMA boxes of (32 dY channels x PX pixels).
// Split-K over pixel slabs across CTAs; partial tiles are accumulated into dW with coalesced
// fp32 reductions (dW must be zeroed by the caller) -- or, in deterministic mode, stored as one
// plane per split and summed in a fixed order by the unpack kernel (bit-repeatable gradients).  Replaces cudnnConvolutionBackwardFilter
// reached from loss.backward() (pretrain_BreastPathQ.py:60).
#pragma once
#include "ptx.cuh"

namespace b2n {

struct WgradParams {
  int M_total;      // N*P*Q pixels of dY
  int P, Q;
  int Cout, Cin, R, S, stride, pad_h, pad_w;
  int Ktot;         // R*S*Cin rows of dW^T
  int num_m_tiles;  // ceil(Ktot / 128)
  int num_n_tiles;  // Cout / BLOCK_N
  int splits;
  int slabs_total;  // ceil(M_total / PX)
  float* dw;        // [Cout][Ktot]; deterministic mode: [splits][Cout][Ktot] partial planes
  int deterministic;  // 0: fp32 reductions into one zeroed plane; 1: every split stores its own
                      // plane (no atomics; b2n_unpack_wgrad sums the planes in a fixed order)
};

constexpr int kWgradThreads = 192;

// PX = pixels per pipeline stage (multiple of 8)
template <int BLOCK_N, int STAGES, int PX>
struct WgradSmem {
  static constexpr int A_BYTES = 128 * PX * 4;
  static constexpr int B_BYTES = BLOCK_N * PX * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = RING_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int BLOCK_N, int STAGES, int PX>
__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_x,
                  const __grid_constant__ CUtensorMap map_dy, const WgradParams p) {
  using L = WgradSmem<BLOCK_N, STAGES, PX>;
  constexpr int CB = 32;                         // channels per TMA box (128 B rows)
  constexpr int A_BOXES = 128 / CB;              // boxes per 128-row tile
  constexpr int BOX_BYTES = PX * 128;
  constexpr int B_BOXES = BLOCK_N / 32;
  constexpr uint32_t ATOM = 4 * 128;             // 4 pixel rows of one 128B/32B-atom swizzle atom
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::RING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int bid = blockIdx.x;
  const int m_tile = bid % p.num_m_tiles; bid /= p.num_m_tiles;
  const int n_tile = bid % p.num_n_tiles; bid /= p.num_n_tiles;
  const int split = bid;
  const int slabs_per = (p.slabs_total + p.splits - 1) / p.splits;
  const int slab_lo = split * slabs_per;
  int slab_hi = slab_lo + slabs_per;
  if (slab_hi > p.slabs_total) slab_hi = p.slabs_total;
  const int num_slabs = slab_hi - slab_lo;  // may be <= 0 for trailing splits

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (num_slabs > 0) {
    if (warp == 0) {
      if (lane == 0) {
        // rows of this tile that exist (last tile of a 9*64-row problem is half empty)
        int valid_boxes = (p.Ktot - m_tile * 128 + CB - 1) / CB;
        if (valid_boxes > A_BOXES) valid_boxes = A_BOXES;
        const uint32_t tx_bytes = valid_boxes * BOX_BYTES + L::B_BYTES;
        // Per-box filter tap / channel offset are slab-invariant; the pixel coordinate is
        // advanced incrementally -- no integer division in the per-stage loop (one elected
        // thread issues everything, so its instruction latency is on the critical path).
        int box_c0[A_BOXES];
        uint16_t box_r[A_BOXES], box_s[A_BOXES];
#pragma unroll
        for (int b = 0; b < A_BOXES; ++b) {
          const int row0 = m_tile * 128 + b * CB;
          const int tap = row0 / p.Cin;
          box_c0[b] = row0 - tap * p.Cin;
          box_r[b] = static_cast<uint16_t>(tap / p.S);
          box_s[b] = static_cast<uint16_t>(tap - (tap / p.S) * p.S);
        }
        const int PQ = p.P * p.Q;
        int m0 = slab_lo * PX;
        int img = m0 / PQ;
        int op = (m0 - img * PQ) / p.Q;
        int oq = m0 - img * PQ - op * p.Q;
        int stage = 0;
        uint32_t phase = 0;
        for (int sl = slab_lo; sl < slab_hi; ++sl) {
          const int base_w = oq * p.stride - p.pad_w;
          const int base_h = op * p.stride - p.pad_h;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
#pragma unroll
          for (int b = 0; b < A_BOXES; ++b) {
            if (b < valid_boxes)
              tma_load_im2col_4d(sa + b * BOX_BYTES, &map_x, &full_bar[stage], box_c0[b], base_w,
                                 base_h, img, box_s[b], box_r[b]);
          }
          // all BLOCK_N / 32 channel groups of dY in one box, group-major in smem
          tma_load_3d(sb, &map_dy, &full_bar[stage], 0, m0, n_tile * B_BOXES);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          m0 += PX;
          oq += PX;
          while (oq >= p.Q) { oq -= p.Q; ++op; }
          while (op >= p.P) { op -= p.P; ++img; }
        }
      }
    } else if (warp == 1) {
      // whole warp, uniform control flow, one elected lane issues; descriptors are one add on
      // a precomputed template (the issuing thread's instruction stream is the critical path)
      constexpr uint32_t idesc = make_idesc_tf32(128, BLOCK_N, 1, 1);
      // MN-major: LBO = stride between 32-channel groups (one TMA box), SBO = stride between
      // 4-pixel groups (one swizzle atom); a K=8 step spans two atoms.
      const uint64_t desc0 = make_smem_desc(0, BOX_BYTES, ATOM, kSwz128B32);
      // (shared high word, low word = stage base + immediate: see umma_tf32_lh)
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32);
      const uint32_t ring_lo = static_cast<uint32_t>(desc0) + (smem_u32(smem) >> 4);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_slabs; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = ring_lo + stage * (L::STAGE_BYTES >> 4);
          const uint32_t b_lo = a_lo + (L::A_BYTES >> 4);
#pragma unroll
          for (int j = 0; j < PX / 8; ++j)
            umma_tf32_lh(tmem_base, a_lo + j * (2 * ATOM >> 4), b_lo + j * (2 * ATOM >> 4), desc_hi,
                         idesc, j != 0 ? 1u : (it != 0 ? 1u : 0u));
          tc_commit(&empty_bar[stage]);
          if (it == num_slabs - 1) tc_commit(tfull_bar);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    } else {
      const int quad = warp & 3;
      const int grow = m_tile * 128 + quad * 32 + lane;  // row of dW^T = tap*Cin + c
      const bool row_ok = grow < p.Ktot;
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
        float v[32];
        tmem_ld_32x32(t_addr + ch * 32, v);
        tmem_ld_wait();
        if (row_ok) {
          float* dst = p.dw + static_cast<size_t>(n_tile * BLOCK_N + ch * 32) * p.Ktot + grow;
          if (p.deterministic) {
            dst += static_cast<size_t>(split) * p.Cout * p.Ktot;
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[static_cast<size_t>(i) * p.Ktot] = v[i];
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(dst + static_cast<size_t>(i) * p.Ktot, v[i]);
          }
        }
      }
    }
  } else if (p.deterministic && warp >= 2) {
    // a trailing split without slabs still owns a plane: it is all zero
    const int grow = m_tile * 128 + (warp & 3) * 32 + lane;
    if (grow < p.Ktot) {
      float* dst = p.dw + static_cast<size_t>(split) * p.Cout * p.Ktot +
                   static_cast<size_t>(n_tile * BLOCK_N) * p.Ktot + grow;
      for (int i = 0; i < BLOCK_N; ++i) dst[static_cast<size_t>(i) * p.Ktot] = 0.f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace b2n
