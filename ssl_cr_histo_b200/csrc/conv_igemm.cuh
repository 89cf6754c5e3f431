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
// SPLIT = true is the error-compensated forward mode: both operands arrive as a (hi, lo) pair
// of FP16 tensors with hi + lo == the FP32 value to ~2^-22 (hi = fp16(v), lo = fp16(v - hi)),
// and each K step issues hi*hi + lo*hi + hi*lo (kind::f16, K = 16 per MMA) into the same FP32
// accumulator -- FP32-grade results at 1.5x the cost of one TF32 pass and the same operand
// bytes.  A single TF32 pass is only good to ~1e-3 on the logits of this network (and flips
// ~0.3 % of the ReLU gates, which the gradients then inherit).  Data-gradient launches
// (SPLIT = false) multiply plain TF32 operands held in FP32 containers.
//
// The same kernel serves forward convs (3x3 s1/s2, 1x1 s2, the space-to-depth stem) and
// data-gradient convs (flipped/transposed weight pack); replaces the cuDNN calls behind
// torchvision BasicBlock.forward (site-packages/torchvision/models/resnet.py:92-100).
#pragma once
#include <cuda_fp16.h>

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
  float* out;             // fp32 result (null when only the FP16 pair is wanted)
  __half* out_h;          // when set the result is (also) stored as a (hi, lo) FP16 pair
  __half* out_l;
  const float* scale;     // per-channel multiplier (eval-mode BN fold) or null
  const float* shift;     // per-channel bias or null
  const float* resid;     // fp32 tensor added in the epilogue or null
  const __half* resid_h;  // (hi, lo) FP16 residual (identity shortcut) or null
  const __half* resid_l;
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
  constexpr int KELEMS = KBYTES / (SPLIT ? 2 : 4);  // fp16 pairs (split) or tf32-in-fp32
  constexpr int MMAS_PER_STAGE = KBYTES / 32;  // one MMA consumes 32 bytes of K (8 tf32 / 16 f16)
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
        // (r, s, channel slice) advance as nested counters: no division per k step -- the
        // single issuing thread's instruction latency is on the critical path.
        int r = 0, s = 0, cs = 0, kcoord = 0;
        for (int ks = 0; ks < num_k_steps; ++ks) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * L::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          tma_load_im2col_4d(st, &map_a, &full_bar[stage], cs * KELEMS, base_w, base_h, img,
                             static_cast<uint16_t>(s), static_cast<uint16_t>(r));
          tma_load_2d(st + OFF_B, &map_b, &full_bar[stage], kcoord, n_tile * BLOCK_N);
          if (SPLIT) {
            tma_load_im2col_4d(st + OFF_A_LO, &map_a_lo, &full_bar[stage], cs * KELEMS, base_w,
                               base_h, img, static_cast<uint16_t>(s), static_cast<uint16_t>(r));
            tma_load_2d(st + OFF_B_LO, &map_b_lo, &full_bar[stage], kcoord, n_tile * BLOCK_N);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          kcoord += KELEMS;
          if (++cs == p.kslices) {
            cs = 0;
            if (++s == p.S) { s = 0; ++r; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc =
          SPLIT ? make_idesc_f16(kBlockM, BLOCK_N) : make_idesc_tf32(kBlockM, BLOCK_N, 0, 0);
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
            if (SPLIT) {
              const uint64_t dal = make_smem_desc(a_addr + OFF_A_LO + j * 32, 16, SBO, SWZ);
              const uint64_t dbl = make_smem_desc(a_addr + OFF_B_LO + j * 32, 16, SBO, SWZ);
              umma_f16(d_tmem, da, db, idesc, (ks | j) != 0 ? 1u : 0u);
              umma_f16(d_tmem, dal, db, idesc, 1u);
              umma_f16(d_tmem, da, dbl, idesc, 1u);
            } else {
              umma_tf32(d_tmem, da, db, idesc, (ks | j) != 0 ? 1u : 0u);
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
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // 8 channels per step
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = v[8 * i + k];
            if (p.scale != nullptr) {
              const float4 s0 = *reinterpret_cast<const float4*>(p.scale + n0 + 8 * i);
              const float4 s1 = *reinterpret_cast<const float4*>(p.scale + n0 + 8 * i + 4);
              o[0] *= s0.x; o[1] *= s0.y; o[2] *= s0.z; o[3] *= s0.w;
              o[4] *= s1.x; o[5] *= s1.y; o[6] *= s1.z; o[7] *= s1.w;
            }
            if (p.shift != nullptr) {
              const float4 s0 = *reinterpret_cast<const float4*>(p.shift + n0 + 8 * i);
              const float4 s1 = *reinterpret_cast<const float4*>(p.shift + n0 + 8 * i + 4);
              o[0] += s0.x; o[1] += s0.y; o[2] += s0.z; o[3] += s0.w;
              o[4] += s1.x; o[5] += s1.y; o[6] += s1.z; o[7] += s1.w;
            }
            if (p.resid != nullptr) {
              const float4 r0 = *reinterpret_cast<const float4*>(p.resid + off + 8 * i);
              const float4 r1 = *reinterpret_cast<const float4*>(p.resid + off + 8 * i + 4);
              float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
              if (p.mask != nullptr) {
                const float4 m0 = *reinterpret_cast<const float4*>(p.mask + off + 8 * i);
                const float4 m1 = *reinterpret_cast<const float4*>(p.mask + off + 8 * i + 4);
                const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) rr[k] = mm[k] > 0.f ? rr[k] : 0.f;
              }
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] += rr[k];
            }
            if (p.resid_h != nullptr) {
              const uint4 rh = *reinterpret_cast<const uint4*>(p.resid_h + off + 8 * i);
              const uint4 rl = *reinterpret_cast<const uint4*>(p.resid_l + off + 8 * i);
              const __half2* h2 = reinterpret_cast<const __half2*>(&rh);
              const __half2* l2 = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 a = __half22float2(h2[k]), b = __half22float2(l2[k]);
                o[2 * k] += a.x + b.x;
                o[2 * k + 1] += a.y + b.y;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] = fmaxf(o[k], 0.f);
            }
            if (p.out_h != nullptr) {
              uint4 ph, pl;
              __half2* h2 = reinterpret_cast<__half2*>(&ph);
              __half2* l2 = reinterpret_cast<__half2*>(&pl);
#pragma unroll
              for (int k = 0; k < 4; ++k) split_f16(o[2 * k], o[2 * k + 1], h2[k], l2[k]);
              *reinterpret_cast<uint4*>(p.out_h + off + 8 * i) = ph;
              *reinterpret_cast<uint4*>(p.out_l + off + 8 * i) = pl;
            }
            if (p.out != nullptr) {
              if (p.round_tf32) {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = tf32_rn(o[k]);
              }
              float4* dst = reinterpret_cast<float4*>(p.out + off + 8 * i);
              dst[0] = make_float4(o[0], o[1], o[2], o[3]);
              dst[1] = make_float4(o[4], o[5], o[6], o[7]);
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
