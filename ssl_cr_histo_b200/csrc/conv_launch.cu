// Host launcher for the tcgen05 implicit-GEMM conv kernel (conv_igemm.cuh).
#include <stdio.h>
#include <stdlib.h>

#include "conv_igemm.cuh"
#include "launch.h"
#include "tmap.h"

namespace b2n {

// Blocks per SM the grid-stride element-wise kernels are capped at (256 threads each; 8 are resident
// at a time, the rest queue behind them).  The weight-gradient kernels run on a side stream beside
// these kernels: their one 192-thread CTA per SM is launched first (trunk.py, ordered overlap) and
// the element-wise blocks fill the remaining thread slots.  Measured with that schedule: 16 beats a
// cap of 7, 8 or 12 (which leave room for the CTA from the start) by 0.3-0.4 ms per step.
int elementwise_blocks_per_sm() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("B2N_EW_BLOCKS_PER_SM");
    v = e != nullptr ? atoi(e) : 16;
    if (v < 1) v = 1;
  }
  return v;
}

int device_sm_count() {
  static int cache[kMaxDevices] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool cached = dev >= 0 && dev < kMaxDevices;
  int sms = cached ? __atomic_load_n(&cache[dev], __ATOMIC_RELAXED) : 0;
  if (sms == 0) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    if (cached) __atomic_store_n(&cache[dev], sms, __ATOMIC_RELAXED);
  }
  return sms;
}

struct ConvMaps {
  CUtensorMap a, b, a_lo, b_lo;
};

constexpr int kMaxDynSmem = 232448;  // 227 KB opt-in limit per CTA on sm_100

// Two epilogue groups per CTA (conv_igemm.cuh) for the launches whose epilogue paces the tile once
// the MMA issuer is out of the way.  B2N_EPI_GROUPS is a bit mask of the kernel families that use
// them (default: both), read per call: 1 = stem forward, 2 = merged stride-2 data gradient.
// (Measured and not kept: layer-1's data gradients gain nothing from a second group -- their tiles
// are bound by the shared-memory port -- and the first block's shortcut + gate + BatchNorm-sum
// epilogue spills at the 168 registers ten warps leave per thread.)  Results are bit-identical
// either way except for the summation order of the per-CTA statistics partials.
static int epi_groups_mask() {
  const char* e = getenv("B2N_EPI_GROUPS");
  return e != nullptr ? atoi(e) : 3;
}

template <int BLOCK_N, int KBYTES, int STAGES, bool SPLIT, bool RES_B, bool HALO = false, int EPI = -1,
          int EPI_WARPS = 4, bool S2M = false, int EPI_GROUPS = 1, int RPS = 1>
static int launch_variant(const ConvMaps& m, const ConvParams& p, int grid, cudaStream_t stream) {
  using L = ConvSmem<S2M ? 4 * BLOCK_N : BLOCK_N, KBYTES, STAGES, SPLIT, RES_B, HALO, EPI_WARPS, EPI_GROUPS, S2M, RPS>;
  auto kern = conv_igemm_kernel<BLOCK_N, KBYTES, STAGES, SPLIT, RES_B, HALO, EPI, EPI_WARPS, S2M, EPI_GROUPS, RPS>;
  if (p.R % RPS != 0) return set_error("conv: %d filter rows per stage do not divide R=%d", RPS, p.R);
  const int smem = L::total(p.R * p.S * p.kslices);
  if (smem > kMaxDynSmem) return set_error("conv: %d B of shared memory needed", smem);
  static PerDeviceMax configured;
  const int want = RES_B ? kMaxDynSmem : L::total(0);
  int dev;
  if (configured.needs(want, &dev)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    if (e != cudaSuccess) return set_error("conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured.set(dev, want);
  }
  kern<<<grid, conv_threads(EPI_WARPS), smem, stream>>>(m.a, m.b, m.a_lo, m.b_lo, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("conv launch: %s", cudaGetErrorString(e));
  return 0;
}

int launch_conv(const ConvArgs& a, cudaStream_t stream) {
  if (a.Cout % 64 != 0) return set_error("conv: Cout=%d must be a multiple of 64", a.Cout);
  const bool split = a.x_h != nullptr;
  // bytes of one K row: all of Cin up to 128 B (narrow rows only exist for HALO-capable convs)
  const int esz = split ? 2 : 4;
  const int kbytes = a.Cin * esz >= 128 ? 128 : a.Cin * esz;
  if (split) {
    if (!a.x_l || !a.w_h || !a.w_l) return set_error("conv: split mode needs x_h, x_l, w_h, w_l");
  } else {
    if (!a.x || !a.w) return set_error("conv: null operand");
  }
  if ((kbytes != 128 && kbytes != 64 && kbytes != 32) || (a.Cin * esz) % kbytes != 0)
    return set_error("conv: unsupported Cin=%d (%d-byte elements)", a.Cin, esz);
  if ((a.out_h != nullptr) != (a.out_l != nullptr) || (a.resid_h != nullptr) != (a.resid_l != nullptr))
    return set_error("conv: FP16 pairs need both planes");
  if (!a.out && !a.out_h) return set_error("conv: no output tensor");
  const TmapDtype dt = split ? kF16 : kF32;
  const int kelems = kbytes / esz;
  const int P = (a.H + a.pad_h_lo + a.pad_h_hi - a.R) / a.stride + 1;
  const int Q = (a.W + a.pad_w_lo + a.pad_w_hi - a.S) / a.stride + 1;
  if (P <= 0 || Q <= 0 || a.N <= 0) return set_error("conv: empty output");
  const long long M = 1ll * a.N * P * Q;
  int block_n = a.Cout % 128 == 0 ? 128 : 64;
  // 256-wide tiles halve the activation-operand smem reads per MAC (the SS-mode UMMA operand
  // fetch, not the math, bounds these kernels) when there are enough tiles to fill the GPU
  if (a.Cout % 256 == 0 && (M / kBlockM) * (a.Cout / 256) >= 2 * device_sm_count()) block_n = 256;
  if (a.force_block_n) block_n = a.force_block_n;
  if (a.Cout % block_n != 0) return set_error("conv: Cout %% BLOCK_N != 0");

  // Tap-sharing HALO variant (conv_igemm.cuh): stride-1 same-width convs whose whole packed
  // weight matrix stays resident in shared memory (layer1's 3x3 convs, the 4x4 s2d stem).
  const int ksteps_full = a.R * a.S * (a.Cin / kelems);
  const bool halo_shape = a.stride == 1 && a.S <= kHaloMaxS && a.pad_w_lo + a.pad_w_hi == a.S - 1 &&
                          block_n == 64 && a.Cout == 64 && a.o_step == 0 && !a.a_tiled2d &&
                          !a.no_resident_weights && !a.no_halo && getenv("B2N_NO_HALO") == nullptr;
  int halo = 0;  // 0 = off, else kbytes of the HALO variant
  if (halo_shape && kbytes == 128 &&
      (split ? ConvSmem<64, 128, 2, true, true, true>::total(ksteps_full)
             : ConvSmem<64, 128, 4, false, true, true>::total(ksteps_full)) <= kMaxDynSmem)
    halo = 128;
  if (halo_shape && kbytes == 32 &&
      (split ? ConvSmem<64, 32, 8, true, true, true>::total(ksteps_full)
             : ConvSmem<64, 32, 8, false, true, true>::total(ksteps_full)) <= kMaxDynSmem)
    halo = 32;
  if (!halo && kbytes == 32) return set_error("conv: 32-byte K rows need the HALO variant");
  if (!halo && kbytes == 64 && !split) return set_error("conv: Cin=16 needs FP16 pair operands");
  const long long M_tiles_over = halo ? 1ll * a.N * P * (Q + a.S - 1) : M;  // raster the M tiles cover
  if (M_tiles_over > 2000000000ll) return set_error("conv: too many output pixels");

  ConvParams p;
  p.M_total = (int)M_tiles_over;
  p.P = P; p.Q = Q;
  p.Cout = a.Cout; p.Cin = a.Cin; p.R = a.R; p.S = a.S;
  p.stride = a.stride; p.pad_h = a.pad_h_lo; p.pad_w = a.pad_w_lo;
  p.num_m_tiles = (int)((M_tiles_over + kBlockM - 1) / kBlockM);
  p.num_n_tiles = a.Cout / block_n;
  p.kslices = a.Cin / kelems;
  p.out = a.out; p.out_h = a.out_h; p.out_l = a.out_l;
  p.scale = a.scale; p.shift = a.shift; p.resid = a.resid; p.resid_h = a.resid_h;
  p.resid_l = a.resid_l; p.mask = a.mask; p.relu = a.relu; p.round_tf32 = a.round_tf32;
  p.stats = a.stats;
  p.gate = a.gate;
  p.bnb_y = a.bnb_y; p.bnb_mean = a.bnb_mean; p.bnb_invstd = a.bnb_invstd;
  p.bnb_scale = a.bnb_scale; p.bnb_shift = a.bnb_shift;
  p.a_lo_nonzero = a.a_lo_nonzero;
  p.a_tiled2d = a.a_tiled2d;
  p.s2m_shortcut = 0;
  p.o_step = a.o_step; p.o_h0 = a.o_h0; p.o_w0 = a.o_w0; p.o_H = a.o_H; p.o_W = a.o_W;
  if (a.o_step != 0 && a.stats != nullptr) return set_error("conv: stats with strided output");
  if ((a.scale != nullptr) != (a.shift != nullptr))
    return set_error("conv: scale and shift come as a pair");
  if (a.mask != nullptr && a.resid == nullptr) return set_error("conv: mask without resid");
  if (a.stats != nullptr && a.bnb_y == nullptr && (a.scale != nullptr || a.shift != nullptr))
    return set_error("conv: batch statistics are taken from the raw output (no scale/shift)");
  if (a.gate != nullptr && a.mask != nullptr) return set_error("conv: gate and mask are exclusive");
  if (a.bnb_y != nullptr) {
    if (!a.stats || !a.bnb_mean || !a.bnb_invstd)
      return set_error("conv: BatchNorm-backward sums need stats, mean and invstd");
    if ((a.bnb_scale != nullptr) != (a.bnb_shift != nullptr))
      return set_error("conv: bnb_scale and bnb_shift come as a pair");
    if (a.resid_h != nullptr || a.o_step != 0 || !a.out || a.out_h)
      return set_error("conv: BatchNorm-backward sums go with a dense fp32 result only");
  } else if (a.bnb_mean || a.bnb_invstd || a.bnb_scale || a.bnb_shift) {
    return set_error("conv: bnb_* given without bnb_y");
  }

  ConvMaps m;
  const uint64_t ktot = (uint64_t)a.R * a.S * a.Cin;
  const void* xa = split ? (const void*)a.x_h : (const void*)a.x;
  const void* wa = split ? (const void*)a.w_h : (const void*)a.w;
  if (a.a_tiled2d) {
    if (split || a.R != 1 || a.S != 1 || a.stride != 1) return set_error("conv: a_tiled2d misuse");
    if (make_tiled_map_2d(&m.a, xa, dt, (uint64_t)M, a.Cin, a.Cin, kBlockM, kelems, kbytes))
      return set_error("conv: %s", tmap_last_error());
  } else if (halo) {
    // S = 1 with the pad columns inside the bounding box: the traversal is the padded-width
    // raster (W + S - 1 positions per row), boxes are kBlockM + S - 1 pixels
    if (make_im2col_map(&m.a, xa, dt, a.N, a.H, a.W, a.Cin, a.R, 1, a.pad_h_lo, a.pad_h_hi, a.pad_w_lo,
                        a.pad_w_hi, 1, kelems, kBlockM + a.S - 1, kbytes))
      return set_error("conv: %s", tmap_last_error());
  } else if (make_im2col_map(&m.a, xa, dt, a.N, a.H, a.W, a.Cin, a.R, a.S, a.pad_h_lo, a.pad_h_hi,
                             a.pad_w_lo, a.pad_w_hi, a.stride, kelems, kBlockM, kbytes))
    return set_error("conv: %s", tmap_last_error());
  if (make_tiled_map_2d(&m.b, wa, dt, a.Cout, ktot, ktot, block_n, kelems, kbytes))
    return set_error("conv: %s", tmap_last_error());
  if (split) {
    if (halo ? make_im2col_map(&m.a_lo, a.x_l, dt, a.N, a.H, a.W, a.Cin, a.R, 1, a.pad_h_lo, a.pad_h_hi,
                               a.pad_w_lo, a.pad_w_hi, 1, kelems, kBlockM + a.S - 1, kbytes)
             : make_im2col_map(&m.a_lo, a.x_l, dt, a.N, a.H, a.W, a.Cin, a.R, a.S, a.pad_h_lo,
                               a.pad_h_hi, a.pad_w_lo, a.pad_w_hi, a.stride, kelems, kBlockM, kbytes))
      return set_error("conv: %s", tmap_last_error());
    if (make_tiled_map_2d(&m.b_lo, a.w_l, dt, a.Cout, ktot, ktot, block_n, kelems, kbytes))
      return set_error("conv: %s", tmap_last_error());
  } else {
    m.a_lo = m.a;
    m.b_lo = m.b;
  }

  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = device_sm_count();
  if (tiles < grid) grid = tiles;

  // epilogue features of this launch; the combinations the network's 64-channel layers use have
  // specialised kernels (see the EPI comment in conv_igemm.cuh), everything else runs generic
  const int epi = (a.stats && !a.bnb_y ? kEpiStats : 0) | (a.scale ? kEpiAffine : 0) |
                  (a.resid ? kEpiResid32 : 0) | (a.mask ? kEpiMask : 0) | (a.resid_h ? kEpiResid16 : 0) |
                  (a.relu ? kEpiRelu : 0) | (a.out ? kEpiOut32 : 0) | (a.out_h ? kEpiOut16 : 0) |
                  (a.round_tf32 ? kEpiRound : 0) | (a.gate ? kEpiGate : 0) | (a.bnb_y ? kEpiBnBwd : 0) |
                  (a.bnb_scale ? kEpiBnGate : 0);
  constexpr int kTrainFwd = kEpiStats | kEpiOut32;                       // raw y + BN statistics
  constexpr int kEvalAct = kEpiAffine | kEpiRelu | kEpiOut16;            // folded BN + ReLU -> pair
  constexpr int kEvalActRes = kEvalAct | kEpiResid16;                    // ... + identity shortcut
  constexpr int kEvalActRes32 = kEvalAct | kEpiResid32;                  // ... + fp32 shortcut (downsample branch)
  constexpr int kEvalStem = kEpiAffine | kEpiRelu | kEpiOut32;
  constexpr int kEvalDown = kEpiAffine | kEpiOut32;                      // folded BN, no ReLU -> fp32
  constexpr int kDgrad = kEpiOut32;
  constexpr int kDgradRes = kEpiOut32 | kEpiResid32;                     // + (pre-gated) shortcut gradient
  constexpr int kDgradResGate = kDgradRes | kEpiGate;                    // ... then the input's ReLU gate
  constexpr int kDgradBn = kEpiOut32 | kEpiBnBwd | kEpiBnGate;           // bn1's gate + backward sums
  constexpr int kDgradResGateBn = kDgradResGate | kEpiBnBwd;             // first block: + the stem BN's sums
  if (halo == 128 && split) {
    if (epi == kTrainFwd) return launch_variant<64, 128, 2, true, true, true, kTrainFwd>(m, p, grid, stream);
    if (epi == kEvalAct) return launch_variant<64, 128, 2, true, true, true, kEvalAct>(m, p, grid, stream);
    if (epi == kEvalActRes) return launch_variant<64, 128, 2, true, true, true, kEvalActRes>(m, p, grid, stream);
    return launch_variant<64, 128, 2, true, true, true>(m, p, grid, stream);
  }
  if (halo == 128) {
    // layer1's data gradients are epilogue-paced (shortcut / gate / BatchNorm-backward operands, 64
    // columns per tile): 8 epilogue warps, one per 32-column chunk and lane quadrant, 3 stages
    if (getenv("B2N_DGRAD_EPI4") == nullptr &&
        ConvSmem<64, 128, 3, false, true, true, 8>::total(ksteps_full) <= kMaxDynSmem) {
      // (measured: 514 -> 448 us and 589 -> 494 us for the shortcut variants; conv2's gradient with
      // the BatchNorm sums has a single operand stream and is better off with 4 warps and 4 stages)
      if (epi == kDgradRes) return launch_variant<64, 128, 3, false, true, true, kDgradRes, 8>(m, p, grid, stream);
      if (epi == kDgradResGate) return launch_variant<64, 128, 3, false, true, true, kDgradResGate, 8>(m, p, grid, stream);
      if (epi == kDgradResGateBn) return launch_variant<64, 128, 3, false, true, true, kDgradResGateBn, 8>(m, p, grid, stream);
    }
    if (epi == kDgrad) return launch_variant<64, 128, 4, false, true, true, kDgrad>(m, p, grid, stream);
    if (epi == kDgradRes) return launch_variant<64, 128, 4, false, true, true, kDgradRes>(m, p, grid, stream);
    if (epi == kDgradResGate) return launch_variant<64, 128, 4, false, true, true, kDgradResGate>(m, p, grid, stream);
    if (epi == kDgradBn) return launch_variant<64, 128, 4, false, true, true, kDgradBn>(m, p, grid, stream);
    if (epi == kDgradResGateBn) return launch_variant<64, 128, 4, false, true, true, kDgradResGateBn>(m, p, grid, stream);
    return launch_variant<64, 128, 4, false, true, true>(m, p, grid, stream);
  }
  if (halo == 32 && split) {  // the stem: epilogue-paced, 8 epilogue warps per tile
    // The 4x4 space-to-depth stem: all four filter rows in one pipeline stage (its 16-channel K rows
    // make a stage of one row only four MMAs, less than the issuer's per-stage overhead), three
    // stages, and -- B2N_EPI_GROUPS bit 1 -- two epilogue groups of eight warps.
    const int rps4 = a.R == 4 && getenv("B2N_STEM_RPS1") == nullptr;
    if (rps4 && (epi_groups_mask() & 1) &&
        ConvSmem<64, 32, 3, true, true, true, 16, 2, false, 4>::total(ksteps_full) <= kMaxDynSmem) {
      if (epi == kTrainFwd) return launch_variant<64, 32, 3, true, true, true, kTrainFwd, 16, false, 2, 4>(m, p, grid, stream);
      if (epi == kEvalStem) return launch_variant<64, 32, 3, true, true, true, kEvalStem, 16, false, 2, 4>(m, p, grid, stream);
    }
    if (rps4 && ConvSmem<64, 32, 3, true, true, true, 8, 1, false, 4>::total(ksteps_full) <= kMaxDynSmem) {
      if (epi == kTrainFwd) return launch_variant<64, 32, 3, true, true, true, kTrainFwd, 8, false, 1, 4>(m, p, grid, stream);
      if (epi == kEvalStem) return launch_variant<64, 32, 3, true, true, true, kEvalStem, 8, false, 1, 4>(m, p, grid, stream);
    }
    if (epi == kTrainFwd) return launch_variant<64, 32, 8, true, true, true, kTrainFwd, 8>(m, p, grid, stream);
    if (epi == kEvalStem) return launch_variant<64, 32, 8, true, true, true, kEvalStem, 8>(m, p, grid, stream);
    return launch_variant<64, 32, 8, true, true, true>(m, p, grid, stream);
  }
  if (halo == 32) return launch_variant<64, 32, 8, false, true, true>(m, p, grid, stream);
  // Resident weights when the whole packed matrix fits next to the activation ring.
  const int ksteps = a.R * a.S * p.kslices;
  if (p.num_n_tiles == 1 && block_n == 64 && !a.no_resident_weights) {
    if (split && kbytes == 128 && ConvSmem<64, 128, 2, true, true>::total(ksteps) <= kMaxDynSmem)
      return launch_variant<64, 128, 2, true, true>(m, p, grid, stream);
    if (split && kbytes == 64 && ConvSmem<64, 64, 5, true, true>::total(ksteps) <= kMaxDynSmem)
      return launch_variant<64, 64, 5, true, true>(m, p, grid, stream);
    if (!split && ConvSmem<64, 128, 4, false, true>::total(ksteps) <= kMaxDynSmem)
      return launch_variant<64, 128, 4, false, true>(m, p, grid, stream);
  }
  if (split && kbytes == 128) {
    if (block_n == 64) return launch_variant<64, 128, 4, true, false>(m, p, grid, stream);
    // layers 2-4, eval mode: folded BN + ReLU (+ identity shortcut) -> FP16 pair with the epilogue
    // compiled in (the shortcut operands of chunk ch + 1 are fetched while chunk ch is processed)
    if (block_n == 128 && epi == kEvalAct) return launch_variant<128, 128, 3, true, false, false, kEvalAct>(m, p, grid, stream);
    if (block_n == 128 && epi == kEvalActRes) return launch_variant<128, 128, 3, true, false, false, kEvalActRes>(m, p, grid, stream);
    if (block_n == 256 && epi == kEvalAct) return launch_variant<256, 128, 2, true, false, false, kEvalAct>(m, p, grid, stream);
    if (block_n == 256 && epi == kEvalActRes) return launch_variant<256, 128, 2, true, false, false, kEvalActRes>(m, p, grid, stream);
    if (block_n == 128 && epi == kTrainFwd) return launch_variant<128, 128, 3, true, false, false, kTrainFwd>(m, p, grid, stream);
    if (block_n == 256 && epi == kTrainFwd) return launch_variant<256, 128, 2, true, false, false, kTrainFwd>(m, p, grid, stream);
    if (block_n == 128 && epi == kEvalActRes32) return launch_variant<128, 128, 3, true, false, false, kEvalActRes32>(m, p, grid, stream);
    if (block_n == 256 && epi == kEvalActRes32) return launch_variant<256, 128, 2, true, false, false, kEvalActRes32>(m, p, grid, stream);
    // eval mode, the 1x1 stride-2 shortcut convs: folded BN only, fp32 result (the block's conv2 adds it)
    if (block_n == 128 && epi == kEvalDown) return launch_variant<128, 128, 3, true, false, false, kEvalDown>(m, p, grid, stream);
    if (block_n == 256 && epi == kEvalDown) return launch_variant<256, 128, 2, true, false, false, kEvalDown>(m, p, grid, stream);
    if (block_n == 128) return launch_variant<128, 128, 3, true, false>(m, p, grid, stream);
    if (block_n == 256) return launch_variant<256, 128, 2, true, false>(m, p, grid, stream);
  } else if (split) {
    if (block_n == 64) return launch_variant<64, 64, 8, true, false>(m, p, grid, stream);
  } else {
    if (block_n == 64) return launch_variant<64, 128, 6, false, false>(m, p, grid, stream);
    // layers 2-4: conv2's data gradient with bn1's gate + BatchNorm-backward sums compiled in
    if (block_n == 128 && epi == kDgradBn)
      return launch_variant<128, 128, 5, false, false, false, kDgradBn>(m, p, grid, stream);
    if (block_n == 256 && epi == kDgradBn)
      return launch_variant<256, 128, 4, false, false, false, kDgradBn>(m, p, grid, stream);
    // ... and conv1's data gradient of layer 2's identity block: shortcut gradient + the input's ReLU
    // gate (400 -> 335 us; the 256-wide tiles of layers 3-4 measure the same compiled or generic)
    if (block_n == 128 && epi == kDgradResGate)
      return launch_variant<128, 128, 5, false, false, false, kDgradResGate>(m, p, grid, stream);
    if (block_n == 128) return launch_variant<128, 128, 5, false, false>(m, p, grid, stream);
    if (block_n == 256) return launch_variant<256, 128, 4, false, false>(m, p, grid, stream);
  }
  return set_error("conv: no kernel variant for BLOCK_N=%d split=%d kbytes=%d", block_n,
                   (int)split, kbytes);
}

int launch_conv_dgrad_s2(const ConvArgs& a, cudaStream_t stream) {
  if (!a.x || !a.w || !a.out) return set_error("conv_dgrad_s2: null tensor");
  if (a.Cin % 32 != 0 || a.Cout % 64 != 0)
    return set_error("conv_dgrad_s2: dY channels %d must be a multiple of 32, dX channels %d of 64", a.Cin, a.Cout);
  if (a.N <= 0 || a.H <= 0 || a.W <= 0 || a.o_H <= 0 || a.o_W <= 0 || (a.o_H + 1) / 2 != a.H ||
      (a.o_W + 1) / 2 != a.W)
    return set_error("conv_dgrad_s2: dY %dx%d does not belong to a stride-2 conv over %dx%d", a.H, a.W, a.o_H, a.o_W);
  const long long M = 1ll * a.N * a.H * a.W;
  if (M > 2000000000ll) return set_error("conv_dgrad_s2: too many pixels");
  ConvParams p = {};
  p.M_total = (int)M;
  p.P = a.H; p.Q = a.W;
  p.Cout = a.Cout; p.Cin = a.Cin; p.R = 2; p.S = 2; p.stride = 1; p.pad_h = 0; p.pad_w = 0;
  p.num_m_tiles = (int)((M + kBlockM - 1) / kBlockM);
  p.num_n_tiles = a.Cout / 64;
  p.kslices = a.Cin / 32;
  p.out = a.out; p.resid = a.resid; p.gate = a.gate;
  p.o_step = 2; p.o_H = a.o_H; p.o_W = a.o_W;
  ConvMaps m;
  // the 2x2 neighbourhood dY[i..i+1, j..j+1]: taps (r, s) in {0,1}^2, one pixel of upper padding
  if (make_im2col_map(&m.a, a.x, kF32, a.N, a.H, a.W, a.Cin, 2, 2, 0, 1, 0, 1, 1, 32, kBlockM, 128))
    return set_error("conv_dgrad_s2: %s", tmap_last_error());
  if (make_tiled_map_2d(&m.b, a.w, kF32, a.Cout, 9ull * a.Cin, 9ull * a.Cin, 64, 32, 128))
    return set_error("conv_dgrad_s2: %s", tmap_last_error());
  m.a_lo = m.a;
  m.b_lo = m.b;
  if (a.x2 != nullptr || a.w2 != nullptr) {
    // fused 1x1 stride-2 shortcut: dY of the shortcut conv (same geometry as dy) and its data-gradient
    // pack [C][K] feed a fifth K block of class (0,0)
    if (a.x2 == nullptr || a.w2 == nullptr) return set_error("conv_dgrad_s2: shortcut needs dy_sc and w_sc");
    if (a.resid != nullptr) return set_error("conv_dgrad_s2: fused shortcut and resid are exclusive");
    if (make_im2col_map(&m.a_lo, a.x2, kF32, a.N, a.H, a.W, a.Cin, 2, 2, 0, 1, 0, 1, 1, 32, kBlockM, 128))
      return set_error("conv_dgrad_s2: %s", tmap_last_error());
    if (make_tiled_map_2d(&m.b_lo, a.w2, kF32, a.Cout, (uint64_t)a.Cin, (uint64_t)a.Cin, 64, 32, 128))
      return set_error("conv_dgrad_s2: %s", tmap_last_error());
    p.s2m_shortcut = 1;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = device_sm_count();
  if (tiles < grid) grid = tiles;
  if (epi_groups_mask() & 2) {
    // two epilogue groups of four warps; the operand combinations the network uses are compiled in
    if (a.resid != nullptr && a.gate != nullptr)
      return launch_variant<64, 128, 4, false, false, false, kEpiOut32 | kEpiResid32 | kEpiGate, 8, true, 2>(m, p, grid, stream);
    if (a.resid == nullptr && a.gate != nullptr)
      return launch_variant<64, 128, 4, false, false, false, kEpiOut32 | kEpiGate, 8, true, 2>(m, p, grid, stream);
    return launch_variant<64, 128, 4, false, false, false, -1, 8, true, 2>(m, p, grid, stream);
  }
  return launch_variant<64, 128, 4, false, false, false, -1, 4, true>(m, p, grid, stream);
}

}  // namespace b2n

// ------------------------------------------------------------------ wgrad
#include "conv_wgrad.cuh"
#include "conv_wgrad_halo.cuh"

namespace b2n {

template <int BLOCK_N, int STAGES, int PX>
static int launch_wgrad_variant(const CUtensorMap& mx, const CUtensorMap& mdy,
                                const WgradParams& p, int grid, cudaStream_t stream) {
  using L = WgradSmem<BLOCK_N, STAGES, PX>;
  auto kern = conv_wgrad_kernel<BLOCK_N, STAGES, PX>;
  static PerDeviceMax configured;
  int dev;
  if (configured.needs(L::TOTAL, &dev)) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return set_error("wgrad: cudaFuncSetAttribute(%d B): %s", L::TOTAL,
                                           cudaGetErrorString(e));
    configured.set(dev, L::TOTAL);
  }
  kern<<<grid, kWgradThreads, L::TOTAL, stream>>>(mx, mdy, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("wgrad launch: %s", cudaGetErrorString(e));
  return 0;
}

template <int NACC, int STAGES>
static int launch_wgrad_halo_variant(const CUtensorMap& mx, const CUtensorMap& mdy,
                                     const WgradHaloParams& p, int grid, cudaStream_t stream) {
  using L = WgradHaloSmem<NACC, STAGES>;
  static_assert(L::TOTAL <= kMaxDynSmem, "wgrad halo stage ring too large");
  auto kern = conv_wgrad_halo_kernel<NACC, STAGES>;
  static PerDeviceMax configured;
  int dev;
  if (configured.needs(L::TOTAL, &dev)) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return set_error("wgrad halo: cudaFuncSetAttribute(%d B): %s", L::TOTAL,
                                           cudaGetErrorString(e));
    configured.set(dev, L::TOTAL);
  }
  kern<<<grid, kWgradHaloThreads, L::TOTAL, stream>>>(mx, mdy, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("wgrad halo launch: %s", cudaGetErrorString(e));
  return 0;
}

// Tap-sharing variant (conv_wgrad_halo.cuh): stride-1 same-width convs with 64 output channels
// whose R * Cin/32 accumulators fit TMEM -- layer1's 3x3 convs (6) and the 4x4 s2d stem (4).
static bool wgrad_is_halo(const WgradArgs& a) {
  return a.stride == 1 && a.Cout == 64 && a.S >= 2 && a.S <= 4 && a.pad_w_lo + a.pad_w_hi == a.S - 1 &&
         a.R * (a.Cin / 32) <= 6 && !a.no_halo && getenv("B2N_NO_HALO") == nullptr;
}
static int wgrad_halo_grid(const WgradArgs& a, int P, int Q) {
  const long long positions = 1ll * a.N * P * (Q + a.S - 1);
  const long long slabs = (positions + kWhPX - 1) / kWhPX;
  long long grid = a.force_splits > 0 ? a.force_splits : device_sm_count();
  return (int)(grid > slabs ? slabs : grid);
}
static int wgrad_generic_splits(const WgradArgs& a, long long M) {
  const int block_n = a.Cout % 256 == 0 ? 256 : (a.Cout % 128 == 0 ? 128 : 64);
  const int px = block_n == 256 ? 32 : 64;
  const long long slabs = (M + px - 1) / px;
  const int out_tiles = ((a.R * a.S * a.Cin + 127) / 128) * (a.Cout / block_n);
  long long splits = a.force_splits > 0 ? a.force_splits : device_sm_count() / out_tiles;
  if (splits < 1) splits = 1;
  return (int)(splits > slabs ? slabs : splits);
}

int wgrad_planes(const WgradArgs& a) {
  if (a.Cin % 32 != 0 || a.Cout % 64 != 0) { set_error("wgrad: unsupported channel counts"); return -1; }
  const int P = (a.H + a.pad_h_lo + a.pad_h_hi - a.R) / a.stride + 1;
  const int Q = (a.W + a.pad_w_lo + a.pad_w_hi - a.S) / a.stride + 1;
  if (P <= 0 || Q <= 0 || a.N <= 0) { set_error("wgrad: empty problem"); return -1; }
  return wgrad_is_halo(a) ? wgrad_halo_grid(a, P, Q) : wgrad_generic_splits(a, 1ll * a.N * P * Q);
}

static int launch_wgrad_halo(const WgradArgs& a, int P, int Q, cudaStream_t stream) {
  WgradHaloParams p;
  p.deterministic = a.deterministic;
  p.P = P; p.Q = Q; p.Cin = a.Cin; p.R = a.R; p.S = a.S;
  p.pad_h = a.pad_h_lo; p.pad_w = a.pad_w_lo;
  p.Ktot = a.R * a.S * a.Cin;
  p.nacc = a.R * (a.Cin / 32);
  const long long positions = 1ll * a.N * P * (Q + a.S - 1);
  if (positions > 2000000000ll) return set_error("wgrad: too many pixels");
  p.slabs_total = (int)((positions + kWhPX - 1) / kWhPX);
  p.dw = a.dw;
  CUtensorMap mx, mdy;
  // X: S = 1 with the pad columns inside the bounding box -> traversal = padded-width raster
  if (make_im2col_map(&mx, a.x, kF32, a.N, a.H, a.W, a.x_channels > 0 ? a.x_channels : a.Cin, a.R, 1,
                      a.pad_h_lo, a.pad_h_hi, a.pad_w_lo, a.pad_w_hi, 1, 32, kWhPX + a.S - 1,
                      kSwizzle128Atom32))
    return set_error("wgrad: %s", tmap_last_error());
  // dY on the same raster: S - 1 zero-filled positions after the last column of every row
  if (make_im2col_map(&mdy, a.dy, kF32, a.N, P, Q, a.Cout, 1, 1, 0, 0, 0, a.S - 1, 1, 32, kWhPX,
                      kSwizzle128Atom32))
    return set_error("wgrad: %s", tmap_last_error());
  const int grid = wgrad_halo_grid(a, P, Q);
  if (p.nacc <= 4) return launch_wgrad_halo_variant<4, 4>(mx, mdy, p, grid, stream);
  return launch_wgrad_halo_variant<6, 3>(mx, mdy, p, grid, stream);
}

int launch_wgrad(const WgradArgs& a, cudaStream_t stream) {
  if (a.Cin % 32 != 0) return set_error("wgrad: Cin=%d must be a multiple of 32", a.Cin);
  if (a.x_channels != 0 && (a.x_channels < 0 || a.x_channels > a.Cin || a.x_channels % 4 != 0 || a.Cin != 32))
    return set_error("wgrad: x_channels=%d needs Cin = 32 and a multiple of 4 channels (16-byte pixel pitch)",
                     a.x_channels);
  if (a.Cout % 64 != 0) return set_error("wgrad: Cout=%d must be a multiple of 64", a.Cout);
  const int P = (a.H + a.pad_h_lo + a.pad_h_hi - a.R) / a.stride + 1;
  const int Q = (a.W + a.pad_w_lo + a.pad_w_hi - a.S) / a.stride + 1;
  if (P <= 0 || Q <= 0 || a.N <= 0) return set_error("wgrad: empty problem");
  const long long M = 1ll * a.N * P * Q;
  if (M > 2000000000ll) return set_error("wgrad: too many pixels");
  if (wgrad_is_halo(a)) return launch_wgrad_halo(a, P, Q, stream);
  const int block_n = a.Cout % 256 == 0 ? 256 : (a.Cout % 128 == 0 ? 128 : 64);

  WgradParams p;
  p.M_total = (int)M; p.P = P; p.Q = Q;
  p.Cout = a.Cout; p.Cin = a.Cin; p.R = a.R; p.S = a.S; p.stride = a.stride;
  p.pad_h = a.pad_h_lo; p.pad_w = a.pad_w_lo;
  p.Ktot = a.R * a.S * a.Cin;
  p.num_m_tiles = (p.Ktot + 127) / 128;
  p.num_n_tiles = a.Cout / block_n;
  const int px = block_n == 256 ? 32 : 64;  // pixels per stage (bigger boxes amortise TMA issue)
  p.slabs_total = (int)((M + px - 1) / px);
  const int out_tiles = p.num_m_tiles * p.num_n_tiles;
  const int splits = wgrad_generic_splits(a, M);
  p.splits = splits;
  p.dw = a.dw;
  p.deterministic = a.deterministic;

  CUtensorMap mx, mdy;
  if (make_im2col_map(&mx, a.x, kF32, a.N, a.H, a.W, a.x_channels > 0 ? a.x_channels : a.Cin, a.R, a.S,
                      a.pad_h_lo, a.pad_h_hi, a.pad_w_lo, a.pad_w_hi, a.stride, 32, px, kSwizzle128Atom32))
    return set_error("wgrad: %s", tmap_last_error());
  // dY as one 3-D box per stage: (32 channels) x (PX pixels) x (BLOCK_N / 32 channel groups)
  if (make_grouped_map_3d(&mdy, a.dy, (uint64_t)M, a.Cout, px, block_n / 32,
                          kSwizzle128Atom32))
    return set_error("wgrad: %s", tmap_last_error());

  const int grid = out_tiles * splits;
  if (block_n == 64) return launch_wgrad_variant<64, 4, 64>(mx, mdy, p, grid, stream);
  if (block_n == 128) return launch_wgrad_variant<128, 3, 64>(mx, mdy, p, grid, stream);
  return launch_wgrad_variant<256, 4, 32>(mx, mdy, p, grid, stream);
}

}  // namespace b2n
