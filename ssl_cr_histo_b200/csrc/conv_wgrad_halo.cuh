// Tap-sharing weight-gradient kernel for the 64-output-channel layers (layer1's 3x3 convs, the
// 4x4 space-to-depth stem) on tcgen05 (sm_100a), TF32 operands / FP32 accumulate.
//
//   dW[k][(r*S+s)*Cin + c] = sum_{pixels m} X[pixel(m) + (r,s)][c] * dY[m][k]
//
// The generic kernel (conv_wgrad.cuh) pulls one activation box per filter tap and re-reads dY
// once per 128-row tile of dW; with 64 output channels that is ~14 bytes from L2 per 64 MACs and
// the L2->SM path, not the tensor pipe, sets its speed.  Here one CTA owns ALL of dW for a slab
// of pixels:
//   * a slab is PX consecutive positions of the padded-width raster (W + S - 1 positions per
//     image row; the pad columns are inside the im2col maps' bounding boxes, so TMA zero-fills
//     X there and dY beyond its last column);
//   * per filter row r and 32-channel group one X box of PX + S - 1 pixels is loaded; the S
//     horizontal taps are the same box shifted by whole pixel rows (128 B).  Both operands are
//     MN-major, so the shift is simply the UMMA descriptor's leading-dimension stride: the
//     M = 128 rows of one MMA are 4 tap slots x 32 channels with LBO = 128 B (slot 3 is idle
//     for a 3-tap filter -- its accumulator rows are dropped);
//   * one accumulator per (r, channel group): R * Cin/32 accumulators of 128 x 64 in TMEM;
//   * split-K over slabs across the persistent CTAs, partial sums reduced into dW with
//     coalesced fp32 atomics (dW zeroed by the caller).
// Replaces cudnnConvolutionBackwardFilter reached from loss.backward() (pretrain_BreastPathQ.py:60).
#pragma once
#include "ptx.cuh"

namespace b2n {

struct WgradHaloParams {
  int P, Q;            // output (= dY) spatial extent; stride 1 "same width": Q == W
  int Cin, R, S;       // Cout is fixed at 64
  int pad_h, pad_w;    // lower padding
  int Ktot;            // R*S*Cin
  int nacc;            // R * Cin/32 accumulators
  int slabs_total;     // ceil(N*P*(Q+S-1) / PX)
  float* dw;           // [64][Ktot]; deterministic mode: [gridDim.x][64][Ktot] partial planes
  int deterministic;   // see WgradParams
};

constexpr int kWgradHaloThreads = 192;
constexpr int kWhPX = 64;                                          // pixels per slab
constexpr int kWhXBox = ((kWhPX + 3) * 128 + 1023) / 1024 * 1024;  // one X box (<= PX+3 pixel rows)
constexpr int kWhYBox = kWhPX * 128;                               // one 32-channel dY box

// NACC = number of accumulators the stage layout is sized for
template <int NACC, int STAGES>
struct WgradHaloSmem {
  static constexpr int STAGE_BYTES = NACC * kWhXBox + 2 * kWhYBox;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = RING_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int NACC, int STAGES>
__global__ void __launch_bounds__(kWgradHaloThreads, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_x,
                       const __grid_constant__ CUtensorMap map_dy, const WgradHaloParams p) {
  using L = WgradHaloSmem<NACC, STAGES>;
  constexpr uint32_t ATOM = 4 * 128;  // 4 pixel rows of one 128B-span / 32B-atom swizzle atom
  constexpr uint32_t TMEM_COLS = NACC * 64 <= 256 ? 256 : 512;
  static_assert(NACC * 64 <= 512, "accumulators must fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::RING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // contiguous slab range of this CTA
  const int slabs_per = (p.slabs_total + gridDim.x - 1) / gridDim.x;
  const int slab_lo = blockIdx.x * slabs_per;
  int slab_hi = slab_lo + slabs_per;
  if (slab_hi > p.slabs_total) slab_hi = p.slabs_total;
  const int num_slabs = slab_hi - slab_lo;
  const int groups = p.Cin >> 5;  // 32-channel groups

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
        const int Wp = p.Q + p.S - 1;
        const int PWp = p.P * Wp;
        const uint32_t tx_bytes =
            static_cast<uint32_t>(p.nacc) * (kWhPX + p.S - 1) * 128 + 2 * kWhYBox;
        int t0 = slab_lo * kWhPX;
        int img = t0 / PWp;
        int op = (t0 - img * PWp) / Wp;
        int oq = t0 - img * PWp - op * Wp;
        int stage = 0;
        uint32_t phase = 0;
        for (int sl = slab_lo; sl < slab_hi; ++sl) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * L::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          int a = 0;
          for (int r = 0; r < p.R; ++r)
            for (int g = 0; g < groups; ++g, ++a)
              tma_load_im2col_4d(st + a * kWhXBox, &map_x, &full_bar[stage], g * 32, oq - p.pad_w,
                                 op - p.pad_h, img, 0, static_cast<uint16_t>(r));
          uint8_t* sb = st + NACC * kWhXBox;
          tma_load_im2col_4d(sb, &map_dy, &full_bar[stage], 0, oq, op, img, 0, 0);
          tma_load_im2col_4d(sb + kWhYBox, &map_dy, &full_bar[stage], 32, oq, op, img, 0, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          oq += kWhPX;
          while (oq >= Wp) { oq -= Wp; ++op; }
          while (op >= p.P) { op -= p.P; ++img; }
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = make_idesc_tf32(128, 64, 1, 1);
      // A: MN-major, the four 32-channel "groups" of the 128 rows are the four tap slots of one
      // box: leading-dimension stride = one pixel row (128 B); 4-pixel atoms are 512 B apart.
      const uint64_t desc_a0 = make_smem_desc(0, 128, ATOM, kSwz128B32);
      // B: two 32-channel dY boxes, group stride = one box
      const uint64_t desc_b0 = make_smem_desc(0, kWhYBox, ATOM, kSwz128B32);
      // (both descriptors share their high word -- SBO and layout; the leading-dimension strides
      // live in the low words -- and a descriptor is the stage base + an immediate: see umma_tf32_lh)
      const uint32_t desc_hi = static_cast<uint32_t>(desc_a0 >> 32);
      const uint32_t ring16 = smem_u32(smem) >> 4;
      const uint32_t a_ring = static_cast<uint32_t>(desc_a0) + ring16;
      const uint32_t b_ring = static_cast<uint32_t>(desc_b0) + ring16 + ((NACC * kWhXBox) >> 4);
      const int nacc = p.nacc;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_slabs; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_ring + stage * (L::STAGE_BYTES >> 4);
          const uint32_t b_lo = b_ring + stage * (L::STAGE_BYTES >> 4);
          const uint32_t acc0 = it != 0 ? 1u : 0u;
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
            if (a < nacc) {
#pragma unroll
              for (int j = 0; j < kWhPX / 8; ++j)
                umma_tf32_lh(tmem_base + a * 64, a_lo + a * (kWhXBox >> 4) + j * (2 * ATOM >> 4),
                             b_lo + j * (2 * ATOM >> 4), desc_hi, idesc, j != 0 ? 1u : acc0);
            }
          }
          tc_commit(&empty_bar[stage]);
          if (it == num_slabs - 1) tc_commit(tfull_bar);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    } else {
      // warp's TMEM lane quadrant == tap slot s; lane == channel within the 32-channel group
      const int slot = warp & 3;
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      if (slot < p.S) {
        for (int a = 0; a < p.nacc; ++a) {
          const int r = a / groups, g = a - r * groups;
          const int grow = (r * p.S + slot) * p.Cin + g * 32 + lane;  // row of dW^T
          const uint32_t t_addr = tmem_base + a * 64 + (static_cast<uint32_t>(slot * 32) << 16);
#pragma unroll 1
          for (int ch = 0; ch < 2; ++ch) {
            float v[32];
            tmem_ld_32x32(t_addr + ch * 32, v);
            tmem_ld_wait();
            float* dst = p.dw + static_cast<size_t>(ch * 32) * p.Ktot + grow;
            if (p.deterministic) {
              dst += static_cast<size_t>(blockIdx.x) * 64 * p.Ktot;
#pragma unroll
              for (int i = 0; i < 32; ++i) dst[static_cast<size_t>(i) * p.Ktot] = v[i];
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) atomicAdd(dst + static_cast<size_t>(i) * p.Ktot, v[i]);
            }
          }
        }
      }
    }
  } else if (p.deterministic && warp >= 2) {
    // a trailing CTA without slabs still owns a plane: it is all zero
    const int slot = warp & 3;
    if (slot < p.S) {
      for (int a = 0; a < p.nacc; ++a) {
        const int r = a / groups, g = a - r * groups;
        float* dst = p.dw + static_cast<size_t>(blockIdx.x) * 64 * p.Ktot +
                     (r * p.S + slot) * p.Cin + g * 32 + lane;
        for (int i = 0; i < 64; ++i) dst[static_cast<size_t>(i) * p.Ktot] = 0.f;
      }
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
