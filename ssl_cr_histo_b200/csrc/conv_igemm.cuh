// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC fp32 storage,
// TF32 operands, FP32 accumulation in TMEM.
//
//   D[m][n] = sum_{r,s,c} X[pixel(m) + (r,s)][c] * Wp[n][(r*S+s)*Cin + c]
//
//   * A operand (activations): im2col-mode TMA, one (128 pixel x KELEMS channel) box per
//     filter tap and channel slice -- no im2col matrix is ever materialised.
//   * B operand (packed weights, K-major [Cout][R*S*Cin]): tiled-mode TMA.
//   * persistent CTAs, static round-robin tile schedule, STAGES-deep smem ring,
//     two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
//
// SPLIT = true is the error-compensated forward mode ("3xTF32"): both operands arrive as a
// (hi, lo) pair of TF32 tensors with hi + lo == the FP32 value to ~2^-22, and each K step
// issues hi*hi + lo*hi + hi*lo into the same accumulator.  A single TF32 pass is only good to
// ~1e-3 on the logits of this network (and flips ~0.3 % of the ReLU gates, which the gradients
// then inherit); the split restores FP32-grade forward results.  Data-gradient launches
// (SPLIT = false) multiply plain TF32 operands.
//
// The same kernel serves forward convs (3x3 s1/s2, 1x1 s2, the space-to-depth stem) and
// data-gradient convs (flipped/transposed weight pack); replaces the cuDNN calls behind
// torchvision BasicBlock.forward (site-packages/torchvision/models/resnet.py:92-100).
#pragma once
#include "ptx.cuh"

namespace b2n {

struct ConvParams {
  int M_total;  // N * P * Q output pixels
  int P, Q;     // output spatial extent
  int Cout;     // output channels == row pitch of out/resid/mask
  int Cin;      // input channels per tap
  int R, S;     // filter taps
  int stride;   // traversal stride
  int pad_h, pad_w;  // lower padding
  int num_m_tiles, num_n_tiles;
  int kslices;  // Cin / KELEMS
  float* out;
  float* out_lo;          // when set the result is stored as a (hi, lo) TF32 pair
  const float* scale;     // per-channel multiplier (eval-mode BN fold) or null
  const float* shift;     // per-channel bias or null
  const float* resid;     // tensor added in the epilogue or null
  const float* resid_lo;  // low part of a split residual or null
  const float* mask;      // when set, resid is only added where mask > 0 (ReLU gate)
  int relu;
  int round_tf32;
  double* stats;  // [2][Cout] per-channel sum / sum of squares of the raw accumulator, or null
};

constexpr int kConvThreads = 192;
constexpr int kBlockM = 128;

template <int BLOCK_N, int KBYTES, int STAGES, bool SPLIT>
struct ConvSmem {
  static constexpr int A_BYTES = kBlockM * KBYTES;
  static constexpr int B_BYTES = BLOCK_N * KBYTES;
  static constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int STATS_FLOATS = 2 * 512;
  // ring | stats accumulators | barriers | tmem ptr   (+1024 slack for manual alignment)
  static constexpr int TOTAL = RING_BYTES + STATS_FLOATS * 4 + (2 * STAGES + 4) * 8 + 16 + 1024;
};

template <int BLOCK_N, int KBYTES, int STAGES, bool SPLIT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a,
                  const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_lo, const ConvParams p) {
  using L = ConvSmem<BLOCK_N, KBYTES, STAGES, SPLIT>;
  constexpr int KELEMS = KBYTES / 4;
  constexpr int MMAS_PER_STAGE = KBYTES / 32;  // tf32: K = 8 elements = 32 bytes per MMA
  constexpr uint32_t SWZ = (KBYTES == 128) ? kSwz128 : (KBYTES == 64 ? kSwz64 : kSwz32);
  constexpr uint32_t SBO = 8 * KBYTES;  // 8 rows of one swizzle atom
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  // stage layout: A_hi | B_hi | A_lo | B_lo
  constexpr int OFF_B = L::A_BYTES;
  constexpr int OFF_A_LO = L::A_BYTES + L::B_BYTES;
  constexpr int OFF_B_LO = 2 * L::A_BYTES + L::B_BYTES;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "BLOCK_N");
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  float* s_stats = reinterpret_cast<float*>(smem + L::RING_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::RING_BYTES + L::STATS_FLOATS * 4);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_k_steps = p.R * p.S * p.kslices;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (SPLIT) {
      tma_prefetch_desc(&map_a_lo);
      tma_prefetch_desc(&map_b_lo);
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < L::STATS_FLOATS; i += kConvThreads) s_stats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int PQ = p.P * p.Q;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.num_n_tiles;
        const int m_tile = tile / p.num_n_tiles;
        const int m0 = m_tile * kBlockM;
        const int img = m0 / PQ;
        const int rem = m0 - img * PQ;
        const int op = rem / p.Q;
        const int oq = rem - op * p.Q;
        const int base_w = oq * p.stride - p.pad_w;
        const int base_h = op * p.stride - p.pad_h;
        for (int ks = 0; ks < num_k_steps; ++ks) {
          const int tap = ks / p.kslices;
          const int cs = ks - tap * p.kslices;
          const int r = tap / p.S;
          const int s = tap - r * p.S;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * L::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          const int kcoord = tap * p.Cin + cs * KELEMS;
          tma_load_im2col_4d(st, &map_a, &full_bar[stage], cs * KELEMS, base_w, base_h, img,
                             static_cast<uint16_t>(s), static_cast<uint16_t>(r));
          tma_load_2d(st + OFF_B, &map_b, &full_bar[stage], kcoord, n_tile * BLOCK_N);
          if (SPLIT) {
            tma_load_im2col_4d(st + OFF_A_LO, &map_a_lo, &full_bar[stage], cs * KELEMS, base_w,
                               base_h, img, static_cast<uint16_t>(s), static_cast<uint16_t>(r));
            tma_load_2d(st + OFF_B_LO, &map_b_lo, &full_bar[stage], kcoord, n_tile * BLOCK_N);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBlockM, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int ks = 0; ks < num_k_steps; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < MMAS_PER_STAGE; ++j) {
            const uint64_t da = make_smem_desc(a_addr + j * 32, 16, SBO, SWZ);
            const uint64_t db = make_smem_desc(a_addr + OFF_B + j * 32, 16, SBO, SWZ);
            umma_tf32(d_tmem, da, db, idesc, (ks | j) != 0 ? 1u : 0u);
            if (SPLIT) {
              const uint64_t dal = make_smem_desc(a_addr + OFF_A_LO + j * 32, 16, SBO, SWZ);
              const uint64_t dbl = make_smem_desc(a_addr + OFF_B_LO + j * 32, 16, SBO, SWZ);
              umma_tf32(d_tmem, dal, db, idesc, 1u);
              umma_tf32(d_tmem, da, dbl, idesc, 1u);
            }
          }
          tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ========================================================= epilogue
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.num_n_tiles;
      const int m_tile = tile / p.num_n_tiles;
      const long long m = static_cast<long long>(m_tile) * kBlockM + row_in_tile;
      const bool row_ok = m < p.M_total;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
        float v[32];
        tmem_ld_32x32(t_addr + ch * 32, v);
        tmem_ld_wait();
        const int n0 = n_tile * BLOCK_N + ch * 32;
        if (p.stats != nullptr) {
          float rs[32], rq[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) { rs[i] = v[i]; rq[i] = v[i] * v[i]; }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float ks_ = up ? rs[i + off] : rs[i];
              const float ss_ = up ? rs[i] : rs[i + off];
              rs[i] = ks_ + __shfl_xor_sync(0xffffffffu, ss_, off);
              const float kq_ = up ? rq[i + off] : rq[i];
              const float sq_ = up ? rq[i] : rq[i + off];
              rq[i] = kq_ + __shfl_xor_sync(0xffffffffu, sq_, off);
            }
          }
          // lane l now owns the 32-row partial sums of column n0 + l
          atomicAdd(&s_stats[n0 + lane], rs[0]);
          atomicAdd(&s_stats[512 + n0 + lane], rq[0]);
        }
        if (row_ok) {
          const size_t off = static_cast<size_t>(m) * p.Cout + n0;
          float4* dst = reinterpret_cast<float4*>(p.out + off);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            if (p.scale != nullptr) {
              const float4 sc = *reinterpret_cast<const float4*>(p.scale + n0 + 4 * i);
              o.x *= sc.x; o.y *= sc.y; o.z *= sc.z; o.w *= sc.w;
            }
            if (p.shift != nullptr) {
              const float4 sh = *reinterpret_cast<const float4*>(p.shift + n0 + 4 * i);
              o.x += sh.x; o.y += sh.y; o.z += sh.z; o.w += sh.w;
            }
            if (p.resid != nullptr) {
              float4 rr = *reinterpret_cast<const float4*>(p.resid + off + 4 * i);
              if (p.resid_lo != nullptr) {
                const float4 rl = *reinterpret_cast<const float4*>(p.resid_lo + off + 4 * i);
                rr.x += rl.x; rr.y += rl.y; rr.z += rl.z; rr.w += rl.w;
              }
              if (p.mask != nullptr) {
                const float4 mk = *reinterpret_cast<const float4*>(p.mask + off + 4 * i);
                rr.x = mk.x > 0.f ? rr.x : 0.f; rr.y = mk.y > 0.f ? rr.y : 0.f;
                rr.z = mk.z > 0.f ? rr.z : 0.f; rr.w = mk.w > 0.f ? rr.w : 0.f;
              }
              o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
            }
            if (p.relu) {
              o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f);
              o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            }
            if (p.out_lo != nullptr) {
              const float4 h = make_float4(tf32_rn(o.x), tf32_rn(o.y), tf32_rn(o.z), tf32_rn(o.w));
              dst[i] = h;
              reinterpret_cast<float4*>(p.out_lo + off)[i] = make_float4(
                  tf32_rn(o.x - h.x), tf32_rn(o.y - h.y), tf32_rn(o.z - h.z), tf32_rn(o.w - h.w));
            } else {
              if (p.round_tf32) {
                o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w);
              }
              dst[i] = o;
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete -> hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.stats != nullptr) {
      // epilogue-only named barrier (warps 2..5 = 128 threads), then flush CTA partials
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = threadIdx.x - 64; c < p.Cout; c += 128) {
        const float a = s_stats[c], b = s_stats[512 + c];
        if (a != 0.f || b != 0.f) {
          atomicAdd(&p.stats[c], static_cast<double>(a));
          atomicAdd(&p.stats[p.Cout + c], static_cast<double>(b));
        }
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
