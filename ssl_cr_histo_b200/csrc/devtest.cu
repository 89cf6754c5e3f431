// Stand-alone bring-up test for the tcgen05 conv kernels (no Python / torch needed).
// Usage: devtest <case>     -- runs one case, prints PASS/FAIL and mismatch diagnostics.
//        devtest list       -- prints the number of cases.
// Inputs are small integers / dyadic fractions, so TF32 products and FP32 sums are exact and
// any mismatch against the CPU loops below is a layout / indexing bug, not rounding.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "launch.h"

using namespace b2n;

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

struct Case {
  const char* name;
  int kind;  // 0 = conv fwd, 1 = wgrad
  int N, H, W, Cin, Cout, R, S, stride, plo, phi;
  int epi;      // conv: 0 plain+stats, 1 scale/shift/relu/round, 2 resid+mask
  int force;    // conv: force BLOCK_N; wgrad: force splits
};

static const Case kCases[] = {
    {"fwd 1x1 tiny (pure GEMM, 1 k-step)", 0, 1, 8, 16, 32, 64, 1, 1, 1, 0, 0, 0, 0},
    {"fwd 1x1 C64 (2 k-slices)", 0, 1, 8, 16, 64, 64, 1, 1, 1, 0, 0, 0, 0},
    {"fwd 3x3 s1 C32->64 one tile", 0, 1, 8, 16, 32, 64, 3, 3, 1, 1, 1, 0, 0},
    {"fwd 3x3 s1 C64->64 56x56 N2 (layer1)", 0, 2, 56, 56, 64, 64, 3, 3, 1, 1, 1, 0, 0},
    {"fwd 3x3 s1 ragged M (N3 7x7 C64->128)", 0, 3, 7, 7, 64, 128, 3, 3, 1, 1, 1, 0, 0},
    {"fwd 3x3 s2 C64->128 56->28 (layer2.0)", 0, 2, 56, 56, 64, 128, 3, 3, 2, 1, 1, 0, 0},
    {"fwd 1x1 s2 C64->128 downsample", 0, 2, 56, 56, 64, 128, 1, 1, 2, 0, 0, 0, 0},
    {"fwd 3x3 s1 C128->256 BLOCK_N=256", 0, 2, 14, 14, 128, 256, 3, 3, 1, 1, 1, 0, 256},
    {"fwd 3x3 s1 C512->512 7x7 N5 (layer4)", 0, 5, 7, 7, 512, 512, 3, 3, 1, 1, 1, 0, 0},
    {"fwd epilogue scale/shift/relu/round", 0, 2, 14, 14, 64, 128, 3, 3, 1, 1, 1, 1, 0},
    {"fwd epilogue resid+mask", 0, 2, 14, 14, 64, 64, 3, 3, 1, 1, 1, 2, 0},
    {"fwd 4x4 s1 pad(2,1) C32->64 (s2d stem)", 0, 2, 20, 20, 32, 64, 4, 4, 1, 2, 1, 0, 0},
    {"fwd many tiles persistent (N8 56x56)", 0, 8, 56, 56, 64, 64, 3, 3, 1, 1, 1, 0, 0},
    {"wgrad 1x1 tiny one slab", 1, 1, 4, 8, 32, 64, 1, 1, 1, 0, 0, 0, 1},
    {"wgrad 1x1 C128->128 splits=1", 1, 1, 8, 16, 128, 128, 1, 1, 1, 0, 0, 0, 1},
    {"wgrad 3x3 s1 C64->64 N2 14x14 splits=1", 1, 2, 14, 14, 64, 64, 3, 3, 1, 1, 1, 0, 1},
    {"wgrad 3x3 s1 C64->64 56x56 N2 auto split", 1, 2, 56, 56, 64, 64, 3, 3, 1, 1, 1, 0, 0},
    {"wgrad 3x3 s2 C64->128 56->28", 1, 2, 56, 56, 64, 128, 3, 3, 2, 1, 1, 0, 0},
    {"wgrad 1x1 s2 C64->128", 1, 2, 56, 56, 64, 128, 1, 1, 2, 0, 0, 0, 0},
    {"wgrad 3x3 C256->512 7x7 N5 ragged", 1, 5, 14, 14, 256, 512, 3, 3, 2, 1, 1, 0, 0},
    {"wgrad 4x4 pad(2,1) C32->64 (s2d stem)", 1, 2, 20, 20, 32, 64, 4, 4, 1, 2, 1, 0, 0},
    // force -1: one box per tap (no HALO)
    {"fwd 3x3 s1 C64->64 56x56 N2 per-tap boxes", 0, 2, 56, 56, 64, 64, 3, 3, 1, 1, 1, 0, -1},
    {"fwd 3x3 s1 C64->64 7x7 N3 HALO ragged", 0, 3, 7, 7, 64, 64, 3, 3, 1, 1, 1, 0, 0},
    {"fwd 3x3 s1 C32->64 13x9 N5 HALO odd sizes", 0, 5, 13, 9, 32, 64, 3, 3, 1, 1, 1, 2, 0},
    {"fwd 3x3 s1 C128->64 14x14 N2 HALO 4 k-slices", 0, 2, 14, 14, 128, 64, 3, 3, 1, 1, 1, 1, 0},
    {"fwd 4x4 s1 pad(2,1) C32->64 HALO (4 taps/box)", 0, 3, 20, 18, 32, 64, 4, 4, 1, 2, 1, 0, 0},
    {"fwd 4x4 s1 pad(2,1) C8->64 HALO 32-byte rows", 0, 3, 20, 18, 8, 64, 4, 4, 1, 2, 1, 0, 0},
    {"fwd 3x3 s1 C8->64 56x56 N2 HALO 32-byte rows", 0, 2, 56, 56, 8, 64, 3, 3, 1, 1, 1, 1, 0},
    {"wgrad 3x3 s1 C64->64 56x56 N2 per-tap boxes", 1, 2, 56, 56, 64, 64, 3, 3, 1, 1, 1, 0, -1},
    {"wgrad 3x3 s1 C64->64 13x9 N5 HALO odd sizes", 1, 5, 13, 9, 64, 64, 3, 3, 1, 1, 1, 0, 0},
    {"wgrad 3x3 s1 C32->64 7x7 N3 HALO one CTA", 1, 3, 7, 7, 32, 64, 3, 3, 1, 1, 1, 0, 1},
    {"wgrad 4x4 pad(2,1) C32->64 20x18 N3 HALO", 1, 3, 20, 18, 32, 64, 4, 4, 1, 2, 1, 0, 0},
};
static const int kNumCases = sizeof(kCases) / sizeof(kCases[0]);

static uint32_t g_seed = 12345;
static int rnd(int lo, int hi) {  // inclusive
  g_seed = g_seed * 1664525u + 1013904223u;
  return lo + (int)((g_seed >> 8) % (uint32_t)(hi - lo + 1));
}

static int report(const char* what, const std::vector<float>& got, const std::vector<double>& ref,
                  int ld, double tol) {
  size_t bad = 0;
  double maxerr = 0;
  for (size_t i = 0; i < got.size(); ++i) {
    double e = fabs((double)got[i] - ref[i]);
    if (!(e <= tol * (1.0 + fabs(ref[i])))) {
      if (bad < 12)
        printf("   mismatch %s[%zu][%zu]: got %.6f expected %.6f\n", what, i / ld, i % ld,
               got[i], ref[i]);
      ++bad;
    }
    if (e > maxerr || e != e) maxerr = e;
  }
  printf("   %s: %zu / %zu mismatches, max abs err %.3e\n", what, bad, got.size(), maxerr);
  if (bad) {
    // row / column histogram of the failures helps identify layout bugs
    size_t rows = got.size() / ld;
    size_t badrows = 0, badcols = 0;
    std::vector<char> rb(rows, 0), cb(ld, 0);
    for (size_t i = 0; i < got.size(); ++i) {
      double e = fabs((double)got[i] - ref[i]);
      if (!(e <= tol * (1.0 + fabs(ref[i])))) { rb[i / ld] = 1; cb[i % ld] = 1; }
    }
    for (size_t r = 0; r < rows; ++r) badrows += rb[r];
    for (int c = 0; c < ld; ++c) badcols += cb[c];
    printf("   bad rows %zu / %zu, bad cols %zu / %d; first bad rows:", badrows, rows, badcols, ld);
    int shown = 0;
    for (size_t r = 0; r < rows && shown < 16; ++r)
      if (rb[r]) { printf(" %zu", r); ++shown; }
    printf("\n");
  }
  return bad ? 1 : 0;
}

static int run_conv(const Case& c) {
  const int P = (c.H + c.plo + c.phi - c.R) / c.stride + 1;
  const int Q = (c.W + c.plo + c.phi - c.S) / c.stride + 1;
  const size_t M = (size_t)c.N * P * Q;
  const int Ktot = c.R * c.S * c.Cin;
  std::vector<float> x((size_t)c.N * c.H * c.W * c.Cin), w((size_t)c.Cout * Ktot);
  for (auto& v : x) v = (float)rnd(-4, 4);
  for (auto& v : w) v = (float)rnd(-2, 2) * 0.25f;
  std::vector<float> scale(c.Cout), shift(c.Cout), resid(M * c.Cout), mask(M * c.Cout);
  for (auto& v : scale) v = (float)rnd(1, 4) * 0.5f;
  for (auto& v : shift) v = (float)rnd(-8, 8);
  for (auto& v : resid) v = (float)rnd(-16, 16);
  for (auto& v : mask) v = (float)rnd(-1, 1);

  std::vector<double> ref(M * c.Cout, 0.0), ssum(c.Cout, 0.0), ssq(c.Cout, 0.0);
  for (int n = 0; n < c.N; ++n)
    for (int p = 0; p < P; ++p)
      for (int q = 0; q < Q; ++q) {
        const size_t m = ((size_t)n * P + p) * Q + q;
        for (int k = 0; k < c.Cout; ++k) {
          double acc = 0;
          for (int r = 0; r < c.R; ++r) {
            const int h = p * c.stride - c.plo + r;
            if (h < 0 || h >= c.H) continue;
            for (int s = 0; s < c.S; ++s) {
              const int ww = q * c.stride - c.plo + s;
              if (ww < 0 || ww >= c.W) continue;
              const float* xp = &x[(((size_t)n * c.H + h) * c.W + ww) * c.Cin];
              const float* wp = &w[(size_t)k * Ktot + (r * c.S + s) * c.Cin];
              for (int ci = 0; ci < c.Cin; ++ci) acc += (double)xp[ci] * wp[ci];
            }
          }
          ssum[k] += acc;
          ssq[k] += acc * acc;
          double o = acc;
          if (c.epi == 1) { o = o * scale[k] + shift[k]; o = o > 0 ? o : 0; }
          if (c.epi == 2) o += mask[m * c.Cout + k] > 0 ? resid[m * c.Cout + k] : 0.0;
          ref[m * c.Cout + k] = o;
        }
      }

  float *dx, *dw, *dout, *dscale, *dshift, *dresid, *dmask;
  double* dstats;
  CK(cudaMalloc(&dx, x.size() * 4));
  CK(cudaMalloc(&dw, w.size() * 4));
  CK(cudaMalloc(&dout, M * c.Cout * 4));
  CK(cudaMalloc(&dscale, c.Cout * 4));
  CK(cudaMalloc(&dshift, c.Cout * 4));
  CK(cudaMalloc(&dresid, M * c.Cout * 4));
  CK(cudaMalloc(&dmask, M * c.Cout * 4));
  CK(cudaMalloc(&dstats, 2 * c.Cout * 8));
  CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dscale, scale.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dshift, shift.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dresid, resid.data(), M * c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dmask, mask.data(), M * c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, M * c.Cout * 4));
  CK(cudaMemset(dstats, 0, 2 * c.Cout * 8));

  ConvArgs a;
  a.x = dx; a.w = dw; a.out = dout;
  a.N = c.N; a.H = c.H; a.W = c.W; a.Cin = c.Cin; a.Cout = c.Cout; a.R = c.R; a.S = c.S;
  a.stride = c.stride;
  a.pad_h_lo = a.pad_w_lo = c.plo; a.pad_h_hi = a.pad_w_hi = c.phi;
  a.force_block_n = c.force > 0 ? c.force : 0;
  a.no_halo = c.force == -1;
  if (c.epi == 0) a.stats = dstats;
  if (c.epi == 1) { a.scale = dscale; a.shift = dshift; a.relu = 1; a.round_tf32 = 1; }
  if (c.epi == 2) { a.resid = dresid; a.mask = dmask; }
  if (launch_conv(a, 0)) { printf("   launch_conv error: %s\n", last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("   kernel failed: %s\n", cudaGetErrorString(e)); return 1; }

  std::vector<float> got(M * c.Cout);
  CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
  int fail = report("out", got, ref, c.Cout, 1e-5);
  if (c.epi == 0) {
    std::vector<double> st(2 * c.Cout);
    CK(cudaMemcpy(st.data(), dstats, st.size() * 8, cudaMemcpyDeviceToHost));
    std::vector<float> g1(c.Cout), g2(c.Cout);
    for (int k = 0; k < c.Cout; ++k) { g1[k] = (float)st[k]; g2[k] = (float)st[c.Cout + k]; }
    fail |= report("sum", g1, ssum, c.Cout, 1e-5);
    fail |= report("sumsq", g2, ssq, c.Cout, 1e-5);
  }
  return fail;
}

static int run_wgrad(const Case& c) {
  const int P = (c.H + c.plo + c.phi - c.R) / c.stride + 1;
  const int Q = (c.W + c.plo + c.phi - c.S) / c.stride + 1;
  const size_t M = (size_t)c.N * P * Q;
  const int Ktot = c.R * c.S * c.Cin;
  std::vector<float> x((size_t)c.N * c.H * c.W * c.Cin), dy(M * c.Cout);
  for (auto& v : x) v = (float)rnd(-2, 2);
  for (auto& v : dy) v = (float)rnd(-2, 2) * 0.5f;
  std::vector<double> ref((size_t)c.Cout * Ktot, 0.0);
  for (int n = 0; n < c.N; ++n)
    for (int p = 0; p < P; ++p)
      for (int q = 0; q < Q; ++q) {
        const size_t m = ((size_t)n * P + p) * Q + q;
        for (int r = 0; r < c.R; ++r) {
          const int h = p * c.stride - c.plo + r;
          if (h < 0 || h >= c.H) continue;
          for (int s = 0; s < c.S; ++s) {
            const int ww = q * c.stride - c.plo + s;
            if (ww < 0 || ww >= c.W) continue;
            const float* xp = &x[(((size_t)n * c.H + h) * c.W + ww) * c.Cin];
            for (int k = 0; k < c.Cout; ++k) {
              const double g = dy[m * c.Cout + k];
              if (g == 0) continue;
              double* rp = &ref[(size_t)k * Ktot + (r * c.S + s) * c.Cin];
              for (int ci = 0; ci < c.Cin; ++ci) rp[ci] += g * xp[ci];
            }
          }
        }
      }
  float *dx, *ddy, *ddw;
  CK(cudaMalloc(&dx, x.size() * 4));
  CK(cudaMalloc(&ddy, dy.size() * 4));
  CK(cudaMalloc(&ddw, ref.size() * 4));
  CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ddy, dy.data(), dy.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(ddw, 0, ref.size() * 4));
  WgradArgs a;
  a.x = dx; a.dy = ddy; a.dw = ddw;
  a.N = c.N; a.H = c.H; a.W = c.W; a.Cin = c.Cin; a.Cout = c.Cout; a.R = c.R; a.S = c.S;
  a.stride = c.stride;
  a.pad_h_lo = a.pad_w_lo = c.plo; a.pad_h_hi = a.pad_w_hi = c.phi;
  a.force_splits = c.force > 0 ? c.force : 0;
  a.no_halo = c.force == -1;
  if (launch_wgrad(a, 0)) { printf("   launch_wgrad error: %s\n", last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("   kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> got(ref.size());
  CK(cudaMemcpy(got.data(), ddw, got.size() * 4, cudaMemcpyDeviceToHost));
  return report("dw", got, ref, Ktot, 1e-5);
}

// timing experiment: 1x1 conv (pure GEMM, M = N*56*56 rows, Cin -> 64) with the A operand
// fetched in im2col mode vs tiled mode -- isolates the TMA mode's per-row cost.
static int run_tma_mode_timing() {
  const int N = 256, H = 56, W = 56, Cout = 64;
  for (int Cin = 64; Cin <= 256; Cin *= 2) {
    const size_t M = (size_t)N * H * W;
    float *dx, *dw, *dout;
    CK(cudaMalloc(&dx, M * Cin * 4));
    CK(cudaMalloc(&dw, (size_t)Cout * Cin * 4));
    CK(cudaMalloc(&dout, M * Cout * 4));
    CK(cudaMemset(dx, 0, M * Cin * 4));
    CK(cudaMemset(dw, 0, (size_t)Cout * Cin * 4));
    for (int mode = 0; mode < 4; ++mode) {
      ConvArgs a;
      a.x = dx; a.w = dw; a.out = dout;
      a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.R = 1; a.S = 1; a.stride = 1;
      a.a_tiled2d = mode & 1;
      a.no_resident_weights = (mode >> 1) & 1;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int it = 0; it < 3; ++it) {
        if (it == 1) cudaEventRecord(e0);
        if (launch_conv(a, 0)) { printf("launch error %s\n", last_error()); return 1; }
      }
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      ms /= 2;
      const double rows = (double)M * (Cin / 32);
      printf("Cin=%3d A-mode=%s weights=%s : %.1f us, %.2f SM-cycles per 128B A row (@1.965GHz, 148 SMs), %.0f GB/s A\n",
             Cin, (mode & 1) ? "tiled " : "im2col", (mode & 2) ? "streamed" : "resident", ms * 1e3,
             ms * 1e-3 * 1.965e9 * 148 / rows, M * Cin * 4.0 / (ms * 1e-3) / 1e9);
    }
    cudaFree(dx); cudaFree(dw); cudaFree(dout);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// TMA delivery-rate microbenchmark: one thread per CTA streams boxes of `rows` x `row_bytes` from
// an L2-resident tensor into a 4-deep shared-memory ring and only waits for their arrival (no
// MMA, no stores).  Compares tiled-mode with im2col-mode boxes: is the im2col unit row-rate bound?
#include "ptx.cuh"
#include "tmap.h"

__global__ void __launch_bounds__(64, 1)
tma_rate_kernel(const __grid_constant__ CUtensorMap map, int im2col, int boxes, int rows,
                int row_bytes, int W, int HW_imgs, int DEPTH, long long* clocks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int box_bytes = rows * row_bytes;
  const int slot_bytes = (box_bytes + 1023) / 1024 * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + DEPTH * slot_bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < boxes + DEPTH; ++i) {
      const int slot = i % DEPTH;
      if (i >= DEPTH) mbar_wait(&bar[slot], ((i / DEPTH) - 1) & 1);
      if (i < boxes) {
        // each CTA walks its own region of the tensor, box after box along the pixel raster
        const int pix = (blockIdx.x * boxes + i) * rows % (HW_imgs - rows - 2 * W);
        mbar_arrive_expect_tx(&bar[slot], box_bytes);
        if (im2col)
          tma_load_im2col_4d(smem + slot * slot_bytes, &map, &bar[slot], 0, pix % W - 1,
                             (pix / W) % W - 1, pix / (W * W), 0, 1);
        else
          tma_load_2d(smem + slot * slot_bytes, &map, &bar[slot], 0, pix);
      }
    }
    clocks[blockIdx.x] = clock64() - t0;
  }
}

// batches of B boxes on one barrier: issue B loads back to back, wait for all, repeat
__global__ void __launch_bounds__(64, 1)
tma_batch_kernel(const __grid_constant__ CUtensorMap map, int im2col, int batches, int B, int rows,
                 int row_bytes, int W, int HW_imgs, long long* clocks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int box_bytes = rows * row_bytes;
  const int slot_bytes = (box_bytes + 1023) / 1024 * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + B * slot_bytes);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < batches; ++i) {
      mbar_arrive_expect_tx(bar, box_bytes * B);
      for (int b = 0; b < B; ++b) {
        const int pix = ((blockIdx.x * batches + i) * B + b) * rows % (HW_imgs - rows - 2 * W);
        if (im2col)
          tma_load_im2col_4d(smem + b * slot_bytes, &map, bar, 0, pix % W - 1, (pix / W) % W - 1,
                             pix / (W * W), 0, 1);
        else
          tma_load_2d(smem + b * slot_bytes, &map, bar, 0, pix);
      }
      mbar_wait(bar, i & 1);
    }
    clocks[blockIdx.x] = clock64() - t0;
  }
}

static int run_tma_batch() {
  const int N = 32, H = 56, W = 56, rows = 130, row_bytes = 128, C = 32;
  const size_t pixels = (size_t)N * H * W;
  float* dx;
  CK(cudaMalloc(&dx, pixels * C * 4));
  CK(cudaMemset(dx, 0, pixels * C * 4));
  long long* dclk;
  CK(cudaMalloc(&dclk, 148 * 8));
  for (int B : {1, 2, 4, 8, 12}) {
    for (int mode = 0; mode < 2; ++mode) {
      CUtensorMap map;
      int rc = mode ? make_im2col_map(&map, dx, kF32, N, H, W, C, 3, 1, 1, 1, 1, 1, 1, C, rows, row_bytes)
                    : make_tiled_map_2d(&map, dx, kF32, pixels, C, C, rows, C, row_bytes);
      if (rc) { printf("map error %s\n", tmap_last_error()); return 1; }
      const int batches = 1000;
      const int smem = B * ((rows * row_bytes + 1023) / 1024 * 1024) + 128 + 1024;
      cudaFuncSetAttribute(tma_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      for (int it = 0; it < 2; ++it)
        tma_batch_kernel<<<148, 64, smem>>>(map, mode, batches, B, rows, row_bytes, W, (int)pixels, dclk);
      CK(cudaDeviceSynchronize());
      long long h[148];
      CK(cudaMemcpy(h, dclk, sizeof h, cudaMemcpyDeviceToHost));
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += (double)h[i];
      avg /= 148;
      printf("batch of %2d boxes (130 x 128 B), %-6s: %7.1f clk/batch  %6.1f clk/box  %5.1f B/clk/SM\n", B,
             mode ? "im2col" : "tiled", avg / batches, avg / batches / B,
             (double)rows * row_bytes * B * batches / avg);
    }
  }
  return 0;
}

static int run_tma_rate() {
  const int N = 32, H = 56, W = 56;  // 32 x 56 x 56 pixels, L2 resident for every row width below
  struct Cfg { int row_bytes, rows; };
  for (Cfg cfg : {Cfg{128, 64}, Cfg{128, 130}, Cfg{32, 131}, Cfg{32, 259}, Cfg{32, 515}, Cfg{32, 67}, Cfg{128, 259}}) {
    const int row_bytes = cfg.row_bytes, rows = cfg.rows;
    const int C = row_bytes / 4;  // fp32 channels per pixel == one K row
    const size_t pixels = (size_t)N * H * W;
    float* dx;
    CK(cudaMalloc(&dx, pixels * C * 4));
    CK(cudaMemset(dx, 0, pixels * C * 4));
    long long* dclk;
    CK(cudaMalloc(&dclk, 148 * 8));
    for (int depth : {1, 2, 4, 8}) {
      for (int mode = (rows > 256 ? 1 : 0); mode < 2; ++mode) {   // tiled boxes hold at most 256 rows
        CUtensorMap map;
        int rc = mode ? make_im2col_map(&map, dx, kF32, N, H, W, C, 3, 1, 1, 1, 1, 1, 1, C, rows, row_bytes)
                      : make_tiled_map_2d(&map, dx, kF32, pixels, C, C, rows, C, row_bytes);
        if (rc) { printf("map error %s\n", tmap_last_error()); return 1; }
        const int boxes = 2000;
        const int smem = depth * ((rows * row_bytes + 1023) / 1024 * 1024) + 128 + 1024;
        cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int it = 0; it < 2; ++it)
          tma_rate_kernel<<<148, 64, smem>>>(map, mode, boxes, rows, row_bytes, W, (int)pixels, depth, dclk);
        CK(cudaDeviceSynchronize());
        long long h[148];
        CK(cudaMemcpy(h, dclk, sizeof h, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += (double)h[i];
        avg /= 148;
        printf("row %3d B, box %3d rows, depth %2d, %-6s: %7.1f clk/box  %5.2f clk/row  %6.1f B/clk/SM\n",
               row_bytes, rows, depth, mode ? "im2col" : "tiled", avg / boxes, avg / boxes / rows,
               (double)rows * row_bytes * boxes / avg);
      }
    }
    cudaFree(dx);
    cudaFree(dclk);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "tmarate")) return run_tma_rate();
  if (argc >= 2 && !strcmp(argv[1], "tmabatch")) return run_tma_batch();
  if (argc >= 2 && !strcmp(argv[1], "tma")) return run_tma_mode_timing();
  if (argc < 2 || !strcmp(argv[1], "list")) { printf("%d\n", kNumCases); return 0; }
  const int id = atoi(argv[1]);
  if (id < 0 || id >= kNumCases) return 3;
  const Case& c = kCases[id];
  printf("[case %d] %s\n", id, c.name);
  const int fail = c.kind == 0 ? run_conv(c) : run_wgrad(c);
  printf("[case %d] %s\n", id, fail ? "FAIL" : "PASS");
  return fail;
}
