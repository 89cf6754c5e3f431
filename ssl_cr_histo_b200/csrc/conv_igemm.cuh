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
// RES_B = true keeps the whole packed weight matrix of the launch resident in shared memory
// (loaded once per CTA) instead of streaming a weight slab with every K step -- possible when
// it fits next to the activation ring (Cout = 64 layers and the stem), where it removes a third
// to two thirds of the L2->SM traffic that bounds this kernel.
//
// HALO = true (stride-1 "same-width" convs with resident weights: layer1's 3x3 and the 4x4
// space-to-depth stem) shares one activation box between the S horizontal filter taps: the M tile
// is 128 consecutive positions of the *padded-width* raster (W + S - 1 positions per image row --
// the im2col map's bounding box spans the pad columns, so TMA zero-fills them), one
// (128 + S - 1) pixel box is loaded per filter row r and channel slice, and tap s is the same box
// read through a descriptor whose start address is advanced by s pixel rows.  A tile then pulls
// R boxes instead of R*S from L2 -- the L2->SM path, not the tensor pipe, bounds the 64-channel
// layers.  The S - 1 raster positions per image row that fall on pad columns produce garbage
// accumulator rows which the epilogue drops.
//
// S2M = true is the merged stride-2 3x3 data gradient: the four output-parity classes of
// dx[2i+ph, 2j+pw] (1, 2, 2 and 4 taps over the 2x2 neighbourhood dY[i..i+1, j..j+1], see pack.cu)
// are four accumulators of ONE tile of 128 dY pixels.  The K loop runs over the four taps x channel
// slices; a stage holds the tap's activation box and the weight slabs of every class that uses the
// tap (4, 2, 2, 1).  The accumulators sit in TMEM in the order (0,0), (0,1), (1,1), (1,0), which
// makes the users of every tap ADJACENT: their slabs are stacked in that order and the tap is ONE
// MMA of width 256 / 128 / 128 / 64 per K step (the activation tile is fetched once per tap, not
// once per class -- the shared-memory port bounds 64-wide TF32 MMAs).  An optional fifth "tap" is
// the block's 1x1 stride-2 shortcut conv: its data gradient only reaches the even-even pixels,
// i.e. class (0,0), so dY_shortcut x W_shortcut is accumulated into that class from a second pair
// of tensor maps (no separate launch, no read-modify-write of dx).  The epilogue then stores the
// four interleaved quarters.  dY is read once instead of once per class and per (class, tap) pair.
//
// The same kernel serves forward convs (3x3 s1/s2, 1x1 s2, the space-to-depth stem) and
// data-gradient convs (flipped/transposed weight pack); replaces the cuDNN calls behind
// torchvision BasicBlock.forward (site-packages/torchvision/models/resnet.py:92-100).
#pragma once
#include <cuda_fp16.h>

#include <type_traits>

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
  const float* gate;      // when set, the whole result is zeroed where gate <= 0 (not with mask)
  // BatchNorm-backward statistics of the result (data-gradient launches): the result g is the
  // gradient w.r.t. relu(bn(y)); with bnb_scale/shift the ReLU gate fmaf(y, scale, shift) > 0 is
  // applied to g first.  stats[0][k] += sum g, stats[1][k] += sum g * (y - mean[k]) * invstd[k].
  const float* bnb_y;
  const float* bnb_mean;
  const float* bnb_invstd;
  const float* bnb_scale;
  const float* bnb_shift;
  int relu;
  int round_tf32;
  double* stats;  // [2][Cout] per-channel sum / sum of squares of the raw accumulator (or the
                  // BatchNorm-backward sums when bnb_y is set), or null
  // Strided output placement (parity classes of a stride-2 data gradient): output pixel (i, j)
  // of image n is stored at (o_h0 + i*o_step, o_w0 + j*o_step) of an [*, o_H, o_W, Cout]
  // tensor; pixels falling outside are dropped.  o_step == 0 selects the dense layout.
  int o_step, o_h0, o_w0, o_H, o_W;
  int s2m_shortcut;         // S2M: a fifth K block -- dY of the 1x1 shortcut conv (map_a_lo) times its
                            // weight pack (map_b_lo) -- is accumulated into class (0,0)
  int a_tiled2d;            // experiment: A is a plain [M][Cin] matrix loaded in tiled mode
  const int* a_lo_nonzero;  // split mode: device flag; 0 => the activation lo plane is all zero
                            // (integer-valued images) and its loads / MMAs are skipped
};

constexpr int kConvThreads = 192;   // producer warp + MMA warp + 4 epilogue warps
// 64-wide tiles whose epilogue, not their MMAs, sets the pace (the stem: K = 256 only) run with 8
// epilogue warps -- two per TMEM lane quadrant, one per 32-column chunk.
__host__ __device__ constexpr int conv_threads(int epi_warps) { return 64 + 32 * epi_warps; }
constexpr int kBlockM = 128;
constexpr int kHaloMaxS = 4;                       // widest filter row the HALO variant handles
constexpr int kHaloRows = kBlockM + kHaloMaxS - 1;  // pixels of the largest HALO activation box

// EPI_WARPS epilogue warps in EPI_GROUPS groups: a group owns one TMEM accumulator stage and takes
// every EPI_GROUPS-th tile of the CTA, EPI_WARPS / EPI_GROUPS warps (4, or 8 for 64-wide tiles)
// share a tile.  NO_STATS: the launch never takes statistics (merged stride-2 data gradient).
// RPS (HALO only): filter rows per pipeline stage -- a stage then holds RPS activation boxes, so a
// launch whose K rows are short (the stem: 16 channels) pays the per-stage barrier / commit
// overhead of its MMA issuer once per tile instead of once per filter row.
template <int BLOCK_N, int KBYTES, int STAGES, bool SPLIT, bool RES_B, bool HALO = false, int EPI_WARPS = 4,
          int EPI_GROUPS = 1, bool NO_STATS = false, int RPS = 1>
struct ConvSmem {
  static constexpr int PLANES = SPLIT ? 2 : 1;
  static constexpr int WPT = EPI_WARPS / EPI_GROUPS;  // epilogue warps per tile
  // HALO boxes hold kBlockM + S - 1 pixels; every plane stays 1024-byte aligned
  static constexpr int A_BYTES = HALO ? (kHaloRows * KBYTES + 1023) / 1024 * 1024 : kBlockM * KBYTES;
  static constexpr int B_BYTES = BLOCK_N * KBYTES;
  static constexpr int A_STAGE = RPS * A_BYTES;   // one plane's activation boxes of a stage
  static constexpr int STAGE_BYTES = PLANES * (A_STAGE + (RES_B ? 0 : B_BYTES));
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  // one private [sum | sumsq] row per epilogue warp; resident-weight launches have one N tile
  static constexpr int STATS_C = NO_STATS ? 0 : (RES_B ? BLOCK_N : 512);
  static constexpr int STATS_FLOATS = EPI_WARPS * 2 * STATS_C;
  // per epilogue warp: SROWS rows x 32 floats.  Two 16-row half rounds where shared memory is
  // scarce; the 8-warp (stem) variant stages all 32 rows at once -- one barrier, twice the ILP
  // (the TF32 resident-weight variant -- layer1's data gradients, 147 KB of weights -- has no room
  // for 8 x 32 staging rows and keeps the two half rounds)
  static constexpr int SROWS = (WPT == 8 && SPLIT && EPI_GROUPS == 1) ? 32 : 16;
  static constexpr int STAGING_BYTES = EPI_WARPS * SROWS * 128;
  // stats | staging | barriers | tmem pointer, rounded up to keep the ring 1024-byte aligned
  static constexpr int CTRL_BYTES =
      (STATS_FLOATS * 4 + STAGING_BYTES + (2 * STAGES + 5) * 8 + 16 + 1023) / 1024 * 1024;
  // [1024-align slack] ctrl | ring | resident B (RES_B only)
  static constexpr int total(int num_k_steps) {
    return 1024 + CTRL_BYTES + RING_BYTES + (RES_B ? num_k_steps * PLANES * B_BYTES : 0);
  }
};

// Epilogue feature bits.  EPI < 0 (generic) tests the ConvParams pointers at run time; EPI >= 0
// compiles exactly that combination -- the 64-channel layers are epilogue-instruction bound (a
// 128 x 64 tile is only ~1-3.5k tensor-pipe cycles), so their launches use specialised kernels.
enum : int {
  kEpiStats = 1, kEpiAffine = 2, kEpiResid32 = 4, kEpiMask = 8, kEpiResid16 = 16, kEpiRelu = 32,
  kEpiOut32 = 64, kEpiOut16 = 128, kEpiRound = 256, kEpiGate = 512, kEpiBnBwd = 1024,
  kEpiBnGate = 2048
};
__host__ __device__ constexpr bool epi_on(int epi, int bit, bool runtime) {
  return epi >= 0 ? (epi & bit) != 0 : runtime;
}

// Merged stride-2 data gradient: which parity classes use tap t = 2*r + s of the 2x2 neighbourhood,
// and where their weight blocks sit in the [C][9*K] pack (block e of class cls, see pack.cu):
//   tap (0,0): classes 0,1,2,3 (blocks 0,1,3,5)   tap (0,1): classes 1,3 (blocks 2,6)
//   tap (1,0): classes 2,3 (blocks 4,7)            tap (1,1): class 3 (block 8)
// Class cls owns accumulator position cls ^ (cls >> 1) (TMEM order 0, 1, 3, 2); a tap's slabs are
// loaded in position order: tap 0 -> blocks 0,1,5,3 (columns 0..255), tap 1 -> 2,6 (columns 64..191),
// tap 2 -> 7,4 (columns 128..255), tap 3 -> 8 (columns 128..191); tap 4 = the shortcut (columns 0..63).
__device__ __forceinline__ int s2m_tap_users(int t) { return t == 0 ? 4 : ((t == 1 || t == 2) ? 2 : 1); }
__device__ __forceinline__ int s2m_block(int t, int i) {
  const unsigned blk_tab[4] = {0x3510u, 0x62u, 0x47u, 0x8u};   // nibble i = block of the tap's i-th slab
  return (blk_tab[t] >> (4 * i)) & 0xF;
}
__host__ __device__ constexpr int s2m_class_pos(int cls) { return cls ^ (cls >> 1); }

template <int BLOCK_N, int KBYTES, int STAGES, bool SPLIT, bool RES_B, bool HALO = false, int EPI = -1,
          int EPI_WARPS = 4, bool S2M = false, int EPI_GROUPS = 1, int RPS = 1>
__global__ void __launch_bounds__(conv_threads(EPI_WARPS), 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a,
                  const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_lo, const ConvParams p) {
  // (S2M: a stage holds one activation box and up to four weight slabs -- one per class using the tap)
  using L = ConvSmem<S2M ? 4 * BLOCK_N : BLOCK_N, KBYTES, STAGES, SPLIT, RES_B, HALO, EPI_WARPS, EPI_GROUPS, S2M, RPS>;
  static_assert(RPS == 1 || HALO, "several filter rows per stage: HALO launches only");
  // Epilogue-paced launches (64-wide tiles: ~0.5-1.7k tensor-pipe cycles against ~500 epilogue
  // instructions per warp and tile, latency-bound at ~0.3 IPC per scheduler) run TWO epilogue
  // groups: group g owns accumulator stage g and takes the CTA's tiles g, g + 2, ..., so two tiles
  // are drained concurrently and the groups' TMEM / shared / global round trips interleave.
  constexpr int WPT = L::WPT;
  static_assert(EPI_GROUPS == 1 || EPI_GROUPS == 2, "one epilogue group per TMEM accumulator stage");
  static_assert(WPT == 4 || (WPT == 8 && BLOCK_N == 64), "8 epilogue warps per tile: one per chunk of a 64-wide tile");
  constexpr int KELEMS = KBYTES / (SPLIT ? 2 : 4);  // fp16 pairs (split) or tf32-in-fp32
  constexpr int MMAS_PER_STAGE = KBYTES / 32;  // one MMA consumes 32 bytes of K (8 tf32 / 16 f16)
  constexpr uint32_t SWZ = (KBYTES == 128) ? kSwz128 : (KBYTES == 64 ? kSwz64 : kSwz32);
  constexpr uint32_t SBO = 8 * KBYTES;  // 8 rows of one swizzle atom
  // STACK: the hi and lo weight tiles sit back to back in shared memory, so hi*hi and hi*lo are ONE
  // MMA of width 2*BLOCK_N over [W_hi ; W_lo] into two accumulator halves (summed by the epilogue)
  // and lo*hi a second one into the first half.  The activation tile is then fetched twice instead
  // of three times per K step -- UMMA operand reads share the SM's 128 B/clk shared-memory port
  // with the TMA writes, and that port, not the tensor pipe, bounds the narrow (N <= 128) tiles.
  constexpr bool STACK = SPLIT && BLOCK_N <= 128;
  static_assert(!S2M || (!SPLIT && !RES_B && !HALO && WPT == 4 && BLOCK_N == 64 &&
                         (EPI < 0 || (EPI & ~(kEpiOut32 | kEpiResid32 | kEpiGate)) == 0)),
                "merged stride-2 data gradient: TF32, streamed weights, 64-wide classes, shortcut / gate epilogue");
  constexpr int NCLS = S2M ? 4 : 1;                        // accumulators (parity classes) per tile
  constexpr int ACC_COLS = S2M ? NCLS * BLOCK_N : (STACK ? 2 * BLOCK_N : BLOCK_N);  // TMEM columns per stage
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
  // streamed stage layout: A_hi | A_lo | B_hi | B_lo   (B part absent when RES_B)
  constexpr int OFF_A_LO = L::A_STAGE;
  constexpr int OFF_B = L::PLANES * L::A_STAGE;
  constexpr int OFF_B_LO = OFF_B + L::B_BYTES;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "BLOCK_N");
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two");
  static_assert(!HALO || RES_B, "HALO needs resident weights");

  extern __shared__ uint8_t smem_raw[];
  // align by offsetting the __shared__ array itself (not through an integer cast) so the compiler
  // keeps the shared address space and emits LDS / STS for the staging tile and the statistics
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* s_stats = reinterpret_cast<float*>(base);
  float4* s_stage = reinterpret_cast<float4*>(base + L::STATS_FLOATS * 4);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + L::STATS_FLOATS * 4 + L::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bres_bar = tempty_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bres_bar + 1);
  uint8_t* smem = base + L::CTRL_BYTES;            // activation (+ weight) ring
  uint8_t* resb = smem + L::RING_BYTES;          // resident weights: [k step][plane][B tile]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_k_steps = S2M ? (4 + p.s2m_shortcut) * p.kslices : p.R * p.S * p.kslices;
  const bool skip_a_lo = SPLIT && p.a_lo_nonzero != nullptr && *p.a_lo_nonzero == 0;

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
      mbar_init(&tempty_bar[i], WPT);  // one arrive per epilogue warp of the group draining the stage
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < L::STATS_FLOATS; i += conv_threads(EPI_WARPS)) s_stats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer
    // (whole warp, uniform control flow, one elected lane issues -- see the MMA warp)
    if (RES_B && elect_one()) {
      // whole weight matrix (num_n_tiles == 1), one barrier for all of it
      mbar_arrive_expect_tx(bres_bar, static_cast<uint32_t>(num_k_steps) * L::PLANES * L::B_BYTES);
      for (int ks = 0; ks < num_k_steps; ++ks) {
        tma_load_2d(resb + ks * L::PLANES * L::B_BYTES, &map_b, bres_bar, ks * KELEMS, 0);
        if (SPLIT)
          tma_load_2d(resb + (ks * L::PLANES + 1) * L::B_BYTES, &map_b_lo, bres_bar, ks * KELEMS, 0);
      }
    }
    __syncwarp();
    const uint32_t a_tx = (HALO ? kBlockM + p.S - 1 : kBlockM) * KBYTES;  // bytes one box delivers
    const uint32_t tx_bytes = L::PLANES * (a_tx + (RES_B ? 0 : L::B_BYTES)) - (skip_a_lo ? a_tx : 0);
    int stage = 0;
    uint32_t phase = 0;
    if (HALO) {
      // one (kBlockM + S - 1)-pixel box per filter row and channel slice; the tile starts at
      // position m0 of the padded-width raster (Q + S - 1 positions per image row)
      const int Wp = p.Q + p.S - 1;
      const int PWp = p.P * Wp;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = tile * kBlockM;
        const int img = m0 / PWp;
        const int rem = m0 - img * PWp;
        const int op = rem / Wp;
        const int base_w = (rem - op * Wp) - p.pad_w;
        const int base_h = op - p.pad_h;
        for (int r0 = 0; r0 < p.R; r0 += RPS) {   // (R is a multiple of RPS: checked by the launcher)
          for (int cs = 0; cs < p.kslices; ++cs) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              uint8_t* st = smem + stage * L::STAGE_BYTES;
              mbar_arrive_expect_tx(&full_bar[stage], RPS * tx_bytes);
#pragma unroll
              for (int i = 0; i < RPS; ++i) {
                tma_load_im2col_4d(st + i * L::A_BYTES, &map_a, &full_bar[stage], cs * KELEMS, base_w, base_h,
                                   img, 0, static_cast<uint16_t>(r0 + i));
                if (SPLIT && !skip_a_lo)
                  tma_load_im2col_4d(st + OFF_A_LO + i * L::A_BYTES, &map_a_lo, &full_bar[stage], cs * KELEMS,
                                     base_w, base_h, img, 0, static_cast<uint16_t>(r0 + i));
              }
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else {
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
        // (r, s, channel slice) advance as nested counters: no division per k step
        int r = 0, s = 0, cs = 0, kcoord = 0;
        for (int ks = 0; ks < num_k_steps; ++ks) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* st = smem + stage * L::STAGE_BYTES;
            if (S2M) {
              // tap (r, s): its box once, then the slabs of the classes that use it, stacked in
              // accumulator order; t == 4 (r == 2): the 1x1 shortcut's dY tile and weight slab
              const int t = 2 * r + s;
              const int users = s2m_tap_users(t);
              constexpr int SLAB = BLOCK_N * KBYTES;
              mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(kBlockM * KBYTES + users * SLAB));
              if (t < 4) {
                tma_load_im2col_4d(st, &map_a, &full_bar[stage], cs * KELEMS, base_w, base_h, img,
                                   static_cast<uint16_t>(s), static_cast<uint16_t>(r));
                for (int i = 0; i < users; ++i)
                  tma_load_2d(st + OFF_B + i * SLAB, &map_b, &full_bar[stage],
                              (s2m_block(t, i) * p.kslices + cs) * KELEMS, n_tile * BLOCK_N);
              } else {
                tma_load_im2col_4d(st, &map_a_lo, &full_bar[stage], cs * KELEMS, base_w, base_h, img, 0, 0);
                tma_load_2d(st + OFF_B, &map_b_lo, &full_bar[stage], cs * KELEMS, n_tile * BLOCK_N);
              }
            } else {
            mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            if (p.a_tiled2d)
              tma_load_2d(st, &map_a, &full_bar[stage], cs * KELEMS, m0);
            else
              tma_load_im2col_4d(st, &map_a, &full_bar[stage], cs * KELEMS, base_w, base_h, img,
                                 static_cast<uint16_t>(s), static_cast<uint16_t>(r));
            if (SPLIT && !skip_a_lo)
              tma_load_im2col_4d(st + OFF_A_LO, &map_a_lo, &full_bar[stage], cs * KELEMS, base_w,
                                 base_h, img, static_cast<uint16_t>(s), static_cast<uint16_t>(r));
            if (!RES_B) {
              tma_load_2d(st + OFF_B, &map_b, &full_bar[stage], kcoord, n_tile * BLOCK_N);
              if (SPLIT)
                tma_load_2d(st + OFF_B_LO, &map_b_lo, &full_bar[stage], kcoord, n_tile * BLOCK_N);
            }
            }
          }
          __syncwarp();
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
    // The whole warp runs this loop with uniform control flow and one elected lane issues: the
    // issuing thread's own instruction stream (descriptor arithmetic between MMAs) is the
    // critical path of a 128 x 64 tile, so everything is kept in warp-uniform values (uniform
    // registers) and a descriptor is one 32-bit add on a precomputed template.
    constexpr uint32_t idesc =
        SPLIT ? make_idesc_f16(kBlockM, BLOCK_N) : make_idesc_tf32(kBlockM, BLOCK_N, 0, 0);
    constexpr uint32_t idesc2 = make_idesc_f16(kBlockM, STACK ? 2 * BLOCK_N : BLOCK_N);
    const uint64_t desc0 = make_smem_desc(0, 16, SBO, SWZ);  // address field filled per MMA
    const uint32_t ring16 = smem_u32(smem) >> 4;             // all offsets in 16-byte units
    const uint32_t resb16 = smem_u32(resb) >> 4;
    if (RES_B) mbar_wait(bres_bar, 0);
    // A descriptor is (shared high word, low word = base + compile-time offset): one 32-bit add.
    const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32);
    const uint32_t desc_lo = static_cast<uint32_t>(desc0);
    constexpr uint32_t PB16 = L::PLANES * L::B_BYTES >> 4;   // one k step of resident weights
    if constexpr (HALO) {
      // The issuing thread's own instruction stream paces these tiles (a 128 x 64 MMA is 32-64
      // tensor-pipe cycles), so the whole tile loop is compiled per filter width S and with the
      // lo-plane MMAs in or out: the taps of a filter row and the MMAs of a K row are unrolled
      // with immediate descriptor offsets, nothing but the barrier waits is decided per stage.
      const int nstages = (p.R / RPS) * p.kslices;
      const uint32_t tap_step = static_cast<uint32_t>(p.kslices) * PB16;   // next tap of the same filter row
      const uint32_t row_step = static_cast<uint32_t>(p.S) * tap_step;     // next filter row
      auto run = [&](auto ns_tag, auto lo_tag) {
        constexpr int NS = decltype(ns_tag)::value;
        constexpr bool WITH_LO = decltype(lo_tag)::value;
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
          uint32_t b_row = desc_lo + resb16;   // weights of the stage's first filter row, tap 0, slice 0
          int cs = 0;
          for (int si = 0; si < nstages; ++si) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_lo = desc_lo + ring16 + stage * (L::STAGE_BYTES >> 4);
              uint32_t b_lo = b_row + cs * PB16;
#pragma unroll
              for (int i = 0; i < RPS; ++i) {
                uint32_t b_tap = b_lo;
#pragma unroll
                for (int sx = 0; sx < NS; ++sx) {
                  // tap sx = the same box, start address advanced by sx pixel rows (KBYTES each).
                  // The 128B swizzle is a function of the absolute smem address (measured: the
                  // shifted descriptor reads correctly with base_offset = 0, not base_offset = sx).
#pragma unroll
                  for (int j = 0; j < MMAS_PER_STAGE; ++j) {
                    const uint32_t da = a_lo + i * (L::A_BYTES >> 4) + (KBYTES / 16) * sx + 2 * j;
                    const uint32_t db = b_tap + 2 * j;
                    const uint32_t accum = (i | sx | j) != 0 ? 1u : (si != 0 ? 1u : 0u);
                    if (STACK) {
                      umma_f16_lh(d_tmem, da, db, desc_hi, idesc2, accum);  // [hi*hi | hi*lo]
                      if (WITH_LO) umma_f16_lh(d_tmem, da + (OFF_A_LO >> 4), db, desc_hi, idesc, 1u);
                    } else {
                      umma_tf32_lh(d_tmem, da, db, desc_hi, idesc, accum);
                    }
                  }
                  b_tap += tap_step;
                }
                b_lo += row_step;
              }
              tc_commit(&empty_bar[stage]);
              if (si == nstages - 1) tc_commit(&tfull_bar[acc]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            if (++cs == p.kslices) { cs = 0; b_row += RPS * row_step; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      };
      using std::integral_constant;
      if (SPLIT && !skip_a_lo) {
        if (p.S == 3) run(integral_constant<int, 3>{}, integral_constant<bool, true>{});
        else if (p.S == 4) run(integral_constant<int, 4>{}, integral_constant<bool, true>{});
        else if (p.S == 2) run(integral_constant<int, 2>{}, integral_constant<bool, true>{});
        else run(integral_constant<int, 1>{}, integral_constant<bool, true>{});
      } else {
        if (p.S == 3) run(integral_constant<int, 3>{}, integral_constant<bool, false>{});
        else if (p.S == 4) run(integral_constant<int, 4>{}, integral_constant<bool, false>{});
        else if (p.S == 2) run(integral_constant<int, 2>{}, integral_constant<bool, false>{});
        else run(integral_constant<int, 1>{}, integral_constant<bool, false>{});
      }
    } else {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        // (descriptors as in the HALO loop: a shared high word, low word = base + immediate)
        // S2M: tap T of one channel slice = one MMA per K step over the stacked slabs of its users
        auto issue_s2m = [&](auto tap_tag, uint32_t a_lo, uint32_t b_lo, bool first) {
          constexpr int T = decltype(tap_tag)::value;
          constexpr uint32_t COL = (T == 0 || T == 4) ? 0 : (T == 1 ? BLOCK_N : 2 * BLOCK_N);   // first accumulator
          constexpr uint32_t WIDTH = T == 0 ? 4 * BLOCK_N : ((T == 1 || T == 2) ? 2 * BLOCK_N : BLOCK_N);
          constexpr uint32_t idesc_t = make_idesc_tf32(kBlockM, WIDTH, 0, 0);
#pragma unroll
          for (int j = 0; j < MMAS_PER_STAGE; ++j)
            umma_tf32_lh(d_tmem + COL, a_lo + 2 * j, b_lo + 2 * j, desc_hi, idesc_t, (first && j == 0) ? 0u : 1u);
        };
        auto issue_plain = [&](auto lo_tag, uint32_t a_lo, uint32_t b_lo, bool first) {
          constexpr bool WITH_LO = decltype(lo_tag)::value;
#pragma unroll
          for (int j = 0; j < MMAS_PER_STAGE; ++j) {
            const uint32_t da = a_lo + 2 * j;
            const uint32_t db = b_lo + 2 * j;
            const uint32_t accum = j != 0 ? 1u : (first ? 0u : 1u);
            if (STACK) {
              umma_f16_lh(d_tmem, da, db, desc_hi, idesc2, accum);  // [hi*hi | hi*lo]
              if (WITH_LO) umma_f16_lh(d_tmem, da + (OFF_A_LO >> 4), db, desc_hi, idesc, 1u);
            } else if (SPLIT) {
              umma_f16_lh(d_tmem, da, db, desc_hi, idesc, accum);
              umma_f16_lh(d_tmem, da, db + (L::B_BYTES >> 4), desc_hi, idesc, 1u);  // lo tile follows hi
              if (WITH_LO) umma_f16_lh(d_tmem, da + (OFF_A_LO >> 4), db, desc_hi, idesc, 1u);
            } else {
              umma_tf32_lh(d_tmem, da, db, desc_hi, idesc, accum);
            }
          }
        };
        int tap_m = 0, cs_m = 0;     // S2M: position in the tap x channel-slice schedule
        uint32_t b_res = desc_lo + resb16;   // RES_B: weights of k step ks
        for (int ks = 0; ks < num_k_steps; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = desc_lo + ring16 + stage * (L::STAGE_BYTES >> 4);
            const uint32_t b_lo = RES_B ? b_res : a_lo + (OFF_B >> 4);
            using std::integral_constant;
            if (S2M) {
              // every class starts with tap (0,0), channel slice 0: that MMA overwrites its accumulator
              const bool first = tap_m == 0 && cs_m == 0;
              if (tap_m == 0) issue_s2m(integral_constant<int, 0>{}, a_lo, b_lo, first);
              else if (tap_m == 1) issue_s2m(integral_constant<int, 1>{}, a_lo, b_lo, false);
              else if (tap_m == 2) issue_s2m(integral_constant<int, 2>{}, a_lo, b_lo, false);
              else if (tap_m == 3) issue_s2m(integral_constant<int, 3>{}, a_lo, b_lo, false);
              else issue_s2m(integral_constant<int, 4>{}, a_lo, b_lo, false);
            } else if (SPLIT && !skip_a_lo) {
              issue_plain(integral_constant<bool, true>{}, a_lo, b_lo, ks == 0);
            } else {
              issue_plain(integral_constant<bool, false>{}, a_lo, b_lo, ks == 0);
            }
            tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (ks == num_k_steps - 1) tc_commit(&tfull_bar[acc]);  // accumulator complete
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (S2M && ++cs_m == p.kslices) { cs_m = 0; ++tap_m; }
          if (RES_B) b_res += PB16;
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ========================================================= epilogue
    const bool f_stats = !S2M && epi_on(EPI, kEpiStats, p.stats != nullptr && p.bnb_y == nullptr);
    const bool f_affine = !S2M && epi_on(EPI, kEpiAffine, p.scale != nullptr);
    const bool f_resid = epi_on(EPI, kEpiResid32, p.resid != nullptr);
    const bool f_mask = !S2M && epi_on(EPI, kEpiMask, p.mask != nullptr);
    const bool f_resid16 = !S2M && epi_on(EPI, kEpiResid16, p.resid_h != nullptr);
    const bool f_relu = !S2M && epi_on(EPI, kEpiRelu, p.relu != 0);
    const bool f_out32 = S2M || epi_on(EPI, kEpiOut32, p.out != nullptr);
    const bool f_out16 = !S2M && epi_on(EPI, kEpiOut16, p.out_h != nullptr);
    const bool f_round = !S2M && epi_on(EPI, kEpiRound, p.round_tf32 != 0);
    const bool f_gate = epi_on(EPI, kEpiGate, p.gate != nullptr);
    const bool f_bnb = !S2M && epi_on(EPI, kEpiBnBwd, p.bnb_y != nullptr);
    const bool f_bngate = !S2M && epi_on(EPI, kEpiBnGate, p.bnb_scale != nullptr);
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int ew = warp - 2;    // epilogue warp index (its private staging tile / statistics row)
    const int row_in_tile = quad * 32 + lane;
    float4* stg = s_stage + ew * (L::SROWS * 8);
    const int grp = EPI_GROUPS == 2 ? ew / WPT : 0;   // epilogue group = the accumulator stage it drains
    int acc = grp;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x + grp * gridDim.x; tile < num_tiles; tile += EPI_GROUPS * gridDim.x) {
      // (resident-weight launches have a single N tile)
      const int n_tile = RES_B ? 0 : tile % p.num_n_tiles;
      const int m_tile = RES_B ? tile : tile / p.num_n_tiles;
      const long long m = static_cast<long long>(m_tile) * kBlockM + row_in_tile;
      // Output rows of this lane's accumulator row, as float4 indices (all-ones = row not stored),
      // redistributed for the transposed stores: after the transpose a lane serves row
      // 4*i + (lane >> 3) of half-round h: slot = 4*h + i.  S2M: per parity class `cls` (its
      // interleaved quarter of the output).
      auto row_ctx = [&](int cls, uint32_t(&rows)[8]) {
        bool row_ok = m < p.M_total;
        size_t row_off = static_cast<size_t>(m) * p.Cout;
        if (S2M) {
          const int PQ = p.P * p.Q;
          const int img = static_cast<int>(m / PQ);
          const int rem = static_cast<int>(m - static_cast<long long>(img) * PQ);
          const int oi = rem / p.Q;
          const int oh = (cls >> 1) + 2 * oi;
          const int ow = (cls & 1) + 2 * (rem - oi * p.Q);
          row_ok = row_ok && oh < p.o_H && ow < p.o_W;
          row_off = ((static_cast<size_t>(img) * p.o_H + oh) * p.o_W + ow) * p.Cout;
        } else if (HALO) {
          // m indexes the padded-width raster: drop the S - 1 pad columns of every image row
          const int Wp = p.Q + p.S - 1;
          const int PWp = p.P * Wp;
          const int img = static_cast<int>(m / PWp);
          const int rem = static_cast<int>(m - static_cast<long long>(img) * PWp);
          const int op = rem / Wp;
          const int oq = rem - op * Wp;
          row_ok = row_ok && oq < p.Q;
          row_off = ((static_cast<size_t>(img) * p.P + op) * p.Q + oq) * p.Cout;
        } else if (p.o_step != 0 && row_ok) {
          const int PQ = p.P * p.Q;
          const int img = static_cast<int>(m / PQ);
          const int rem = static_cast<int>(m - static_cast<long long>(img) * PQ);
          const int oi = rem / p.Q;
          const int oh = p.o_h0 + oi * p.o_step;
          const int ow = p.o_w0 + (rem - oi * p.Q) * p.o_step;
          row_ok = oh < p.o_H && ow < p.o_W;
          row_off = ((static_cast<size_t>(img) * p.o_H + oh) * p.o_W + ow) * p.Cout;
        }
        const uint32_t my_row4 = row_ok ? static_cast<uint32_t>(row_off >> 2) : 0xFFFFFFFFu;
#pragma unroll
        for (int sl = 0; sl < 8; ++sl)
          rows[sl] = __shfl_sync(0xffffffffu, my_row4, (sl >> 2) * 16 + 4 * (sl & 3) + (lane >> 3));
      };
      uint32_t row4[8];
      if (!S2M) row_ctx(0, row4);
      // Residual operands are fetched up front -- eight independent loads per lane and chunk in
      // flight instead of one load -> use -> store round trip per row group (the stores may alias,
      // so the compiler cannot hoist).  64-wide tiles fetch the whole tile *before* waiting for the
      // accumulator, hiding the HBM latency behind the tile's MMAs.
      // pre_m holds the mask or the gate (mutually exclusive), pre_x the FP16 residual pair
      // (hi in .x/.y, lo in .z/.w) or the bits of the BatchNorm-backward y (mutually exclusive).
      // `rows`: row_ctx of the tile (class); `with_resid`: the fp32 shortcut is added to this chunk
      auto prefetch = [&](const uint32_t(&rows)[8], bool with_resid, int ch, float4(&pre_r)[8],
                          float4(&pre_m)[8], uint4(&pre_x)[8]) {
        const int c4 = n_tile * BLOCK_N + ch * 32 + 4 * (lane & 7);
        if (with_resid) {
#pragma unroll
          for (int sl = 0; sl < 8; ++sl)
            pre_r[sl] = rows[sl] != 0xFFFFFFFFu
                            ? *reinterpret_cast<const float4*>(p.resid + (static_cast<size_t>(rows[sl]) << 2) + c4)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if ((with_resid && f_mask) || f_gate) {
          const float* src = f_gate ? p.gate : p.mask;
#pragma unroll
          for (int sl = 0; sl < 8; ++sl)
            pre_m[sl] = rows[sl] != 0xFFFFFFFFu
                            ? *reinterpret_cast<const float4*>(src + (static_cast<size_t>(rows[sl]) << 2) + c4)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (f_resid16) {
#pragma unroll
          for (int sl = 0; sl < 8; ++sl) {
            const bool ok = rows[sl] != 0xFFFFFFFFu;
            const size_t o = (static_cast<size_t>(rows[sl]) << 2) + c4;
            const uint2 h = ok ? *reinterpret_cast<const uint2*>(p.resid_h + o) : make_uint2(0u, 0u);
            const uint2 l = ok ? *reinterpret_cast<const uint2*>(p.resid_l + o) : make_uint2(0u, 0u);
            pre_x[sl] = make_uint4(h.x, h.y, l.x, l.y);
          }
        } else if (f_bnb) {
#pragma unroll
          for (int sl = 0; sl < 8; ++sl)
            pre_x[sl] = rows[sl] != 0xFFFFFFFFu
                            ? *reinterpret_cast<const uint4*>(p.bnb_y + (static_cast<size_t>(rows[sl]) << 2) + c4)
                            : make_uint4(0u, 0u, 0u, 0u);
        }
      };
      const uint32_t t_addr = tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(quad * 32) << 16);
      // `taddr`: TMEM address of the accumulator (class) this chunk belongs to
      auto process = [&](const uint32_t(&rows)[8], bool with_resid, uint32_t taddr, int ch,
                         const float4(&pre_r)[8], const float4(&pre_m)[8], const uint4(&pre_x)[8]) {
        float v[32];
        tmem_ld_32x32(taddr + ch * 32, v);
        const int n0 = n_tile * BLOCK_N + ch * 32;
        const int c4 = n0 + 4 * (lane & 7);
        // per-channel constants of the BatchNorm-backward statistics (this lane's four channels)
        float4 b_mu = make_float4(0.f, 0.f, 0.f, 0.f), b_is = b_mu, b_sc = b_mu, b_sh = b_mu;
        if (f_bnb) {
          b_mu = *reinterpret_cast<const float4*>(p.bnb_mean + c4);
          b_is = *reinterpret_cast<const float4*>(p.bnb_invstd + c4);
          if (f_bngate) {
            b_sc = *reinterpret_cast<const float4*>(p.bnb_scale + c4);
            b_sh = *reinterpret_cast<const float4*>(p.bnb_shift + c4);
          }
        }
        if (STACK) {  // second accumulator half: the hi * W_lo products
          float v2[32];
          tmem_ld_32x32(taddr + BLOCK_N + ch * 32, v2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += v2[i];
        }
        tmem_ld_wait();
        // per-channel affine (eval-mode BatchNorm fold): applied after the transpose, where a lane
        // owns four channels -- eight constants per lane instead of sixty-four
        float a_sc[4] = {1.f, 1.f, 1.f, 1.f}, a_sh[4] = {0.f, 0.f, 0.f, 0.f};
        if (f_affine) {
          const float4 sc = *reinterpret_cast<const float4*>(p.scale + c4);
          const float4 sh = *reinterpret_cast<const float4*>(p.shift + c4);
          a_sc[0] = sc.x; a_sc[1] = sc.y; a_sc[2] = sc.z; a_sc[3] = sc.w;
          a_sh[0] = sh.x; a_sh[1] = sh.y; a_sh[2] = sh.z; a_sh[3] = sh.w;
        }
        // Transpose through the warp's staging tile (16 rows x 128 B per round, 16-byte chunks
        // XOR-swizzled by row) so that global memory is touched row-major: 8 lanes cover the
        // 128 contiguous bytes of one output row -- 4 lines per warp instruction instead of 32.
        float st_s[4] = {0.f, 0.f, 0.f, 0.f}, st_q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 32 / L::SROWS; ++h) {
          if (lane / L::SROWS == h) {
            const int r = lane % L::SROWS;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              stg[r * 8 + (c ^ (r & 7))] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < L::SROWS / 4; ++i) {
            const int r = 4 * i + (lane >> 3);
            const int c = lane & 7;
            const float4 t = stg[r * 8 + (c ^ (r & 7))];
            const int sl = (L::SROWS / 4) * h + i;
            if (rows[sl] != 0xFFFFFFFFu) {
              const size_t off = (static_cast<size_t>(rows[sl]) << 2) + c4;
              float o[4] = {t.x, t.y, t.z, t.w};
              if (f_affine) {
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = o[k] * a_sc[k] + a_sh[k];
              }
              if (f_stats) {  // BN batch statistics of the raw conv output
#pragma unroll
                for (int k = 0; k < 4; ++k) { st_s[k] += o[k]; st_q[k] += o[k] * o[k]; }
              }
              if (with_resid) {
                const float4 r0 = pre_r[sl];
                float rr[4] = {r0.x, r0.y, r0.z, r0.w};
                if (f_mask && !f_gate) {
                  const float4 m0 = pre_m[sl];
                  rr[0] = m0.x > 0.f ? rr[0] : 0.f; rr[1] = m0.y > 0.f ? rr[1] : 0.f;
                  rr[2] = m0.z > 0.f ? rr[2] : 0.f; rr[3] = m0.w > 0.f ? rr[3] : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] += rr[k];
              }
              if (f_resid16) {
                const uint2 rh = make_uint2(pre_x[sl].x, pre_x[sl].y);
                const uint2 rl = make_uint2(pre_x[sl].z, pre_x[sl].w);
                const __half2* h2 = reinterpret_cast<const __half2*>(&rh);
                const __half2* l2 = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  const float2 a = __half22float2(h2[k]), b = __half22float2(l2[k]);
                  o[2 * k] += a.x + b.x;
                  o[2 * k + 1] += a.y + b.y;
                }
              }
              if (f_relu) {
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = fmaxf(o[k], 0.f);
              }
              if (f_gate) {   // ReLU gate of the tensor this gradient belongs to
                const float4 g0 = pre_m[sl];
                o[0] = g0.x > 0.f ? o[0] : 0.f; o[1] = g0.y > 0.f ? o[1] : 0.f;
                o[2] = g0.z > 0.f ? o[2] : 0.f; o[3] = g0.w > 0.f ? o[3] : 0.f;
              }
              if (f_bnb) {    // BatchNorm-backward sums of the (gated) gradient
                const float yv[4] = {__uint_as_float(pre_x[sl].x), __uint_as_float(pre_x[sl].y),
                                     __uint_as_float(pre_x[sl].z), __uint_as_float(pre_x[sl].w)};
                const float mu[4] = {b_mu.x, b_mu.y, b_mu.z, b_mu.w};
                const float is[4] = {b_is.x, b_is.y, b_is.z, b_is.w};
                if (f_bngate) {
                  const float sc[4] = {b_sc.x, b_sc.y, b_sc.z, b_sc.w};
                  const float sh[4] = {b_sh.x, b_sh.y, b_sh.z, b_sh.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) o[k] = fmaf(yv[k], sc[k], sh[k]) > 0.f ? o[k] : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  st_s[k] += o[k];
                  st_q[k] += o[k] * (yv[k] - mu[k]) * is[k];
                }
              }
              if (f_out16) {
                uint2 ph, pl;
                __half2* h2 = reinterpret_cast<__half2*>(&ph);
                __half2* l2 = reinterpret_cast<__half2*>(&pl);
                split_f16(o[0], o[1], h2[0], l2[0]);
                split_f16(o[2], o[3], h2[1], l2[1]);
                *reinterpret_cast<uint2*>(p.out_h + off) = ph;
                *reinterpret_cast<uint2*>(p.out_l + off) = pl;
              }
              if (f_out32) {
                if (f_round) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) o[k] = tf32_rn(o[k]);
                }
                *reinterpret_cast<float4*>(p.out + off) = make_float4(o[0], o[1], o[2], o[3]);
              }
            }
          }
          __syncwarp();
        }
        if (f_stats || f_bnb) {
          // lanes l, l+8, l+16, l+24 hold the same four channels (different rows); each epilogue
          // warp accumulates into its own smem row (no atomics, no cross-warp contention)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            st_s[k] += __shfl_xor_sync(0xffffffffu, st_s[k], 8);
            st_q[k] += __shfl_xor_sync(0xffffffffu, st_q[k], 8);
            st_s[k] += __shfl_xor_sync(0xffffffffu, st_s[k], 16);
            st_q[k] += __shfl_xor_sync(0xffffffffu, st_q[k], 16);
          }
          if (lane < 8) {
            float* mine = s_stats + ew * (2 * L::STATS_C) + n0 + 4 * lane;
#pragma unroll
            for (int k = 0; k < 4; ++k) { mine[k] += st_s[k]; mine[L::STATS_C + k] += st_q[k]; }
          }
        }
      };
      if constexpr (S2M) {
        // four parity classes x two 32-column chunks = eight steps; the gate / shortcut operands of
        // step st + 1 are fetched while step st is processed (the first ones before the accumulator
        // wait), each class with its own output rows.  Only class (0,0) carries the 1x1 shortcut.
        uint32_t rows[2][8];
        float4 pr[2][8], pm[2][8];
        uint4 px[2][8];
        row_ctx(0, rows[0]);
        prefetch(rows[0], f_resid, 0, pr[0], pm[0], px[0]);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll
        for (int st = 0; st < 8; ++st) {
          const int cls = st >> 1, ch = st & 1;
          if (st + 1 < 8) {
            const int ncls = (st + 1) >> 1;
            if (((st + 1) & 1) == 0) row_ctx(ncls, rows[ncls & 1]);
            prefetch(rows[ncls & 1], f_resid && ncls == 0, (st + 1) & 1, pr[(st + 1) & 1], pm[(st + 1) & 1],
                     px[(st + 1) & 1]);
          }
          process(rows[cls & 1], f_resid && cls == 0, t_addr + s2m_class_pos(cls) * BLOCK_N, ch, pr[st & 1],
                  pm[st & 1], px[st & 1]);
        }
      } else if constexpr (WPT == 8) {
        // one chunk per warp: the group's first four warps take columns 0..31, the others 32..63
        const int ch = (ew % WPT) >> 2;
        float4 pr[8], pm[8];
        uint4 px[8];
        prefetch(row4, f_resid, ch, pr, pm, px);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        process(row4, f_resid, t_addr, ch, pr, pm, px);
      } else if constexpr (BLOCK_N == 64 && EPI >= 0) {
        float4 pr0[8], pm0[8], pr1[8], pm1[8];
        uint4 px0[8], px1[8];
        prefetch(row4, f_resid, 0, pr0, pm0, px0);
        prefetch(row4, f_resid, 1, pr1, pm1, px1);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        process(row4, f_resid, t_addr, 0, pr0, pm0, px0);
        process(row4, f_resid, t_addr, 1, pr1, pm1, px1);
      } else if constexpr (EPI >= 0) {
        // wide tiles with a compiled-in epilogue: the operands of chunk 0 are fetched before the
        // accumulator wait and those of chunk ch + 1 while chunk ch is processed
        float4 pr[2][8], pm[2][8];
        uint4 px[2][8];
        prefetch(row4, f_resid, 0, pr[0], pm[0], px[0]);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        // (two chunks per iteration so that the buffer index stays a compile-time constant
        // without unrolling all eight chunks of a 256-wide tile)
#pragma unroll 1
        for (int ch = 0; ch < BLOCK_N / 32; ch += 2) {
          prefetch(row4, f_resid, ch + 1, pr[1], pm[1], px[1]);
          process(row4, f_resid, t_addr, ch, pr[0], pm[0], px[0]);
          if (ch + 2 < BLOCK_N / 32) prefetch(row4, f_resid, ch + 2, pr[0], pm[0], px[0]);
          process(row4, f_resid, t_addr, ch + 1, pr[1], pm[1], px[1]);
        }
      } else {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
          float4 pr[8], pm[8];
          uint4 px[8];
          prefetch(row4, f_resid, ch, pr, pm, px);
          process(row4, f_resid, t_addr, ch, pr, pm, px);
        }
      }
      // all TMEM reads of this accumulator stage are complete -> hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (EPI_GROUPS == 2) {
        acc_phase ^= 1;   // the group's next tile lands in the same stage
      } else if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (f_stats || f_bnb) {
      // epilogue-only named barrier (warps 2..5 = 128 threads), then flush CTA partials
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      for (int c = threadIdx.x - 64; c < p.Cout; c += 32 * EPI_WARPS) {
        constexpr int SC = L::STATS_C;
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < EPI_WARPS; ++w) {
          a += s_stats[2 * w * SC + c];
          b += s_stats[(2 * w + 1) * SC + c];
        }
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
