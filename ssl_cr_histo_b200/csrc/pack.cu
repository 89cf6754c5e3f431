// Weight re-layout kernels.  Parameters stay in PyTorch's (K,C,R,S) layout so reference
// checkpoints load unchanged; the tensor-core kernels consume K-major
// packs with the reduction index (tap, channel) contiguous, values rounded to TF32.
#include <cuda_fp16.h>

#include "launch.h"
#include "ptx.cuh"

namespace b2n {

// forward operand (FP16 hi/lo pair):  wf[k][(r*S+s)*C + c] = w[k][c][r][s]
// (every pack is a grid-stride range [first, total) step `step` so that one launch can run many)
__device__ __forceinline__ void pack_fwd_range(const float* __restrict__ w, __half* __restrict__ wf_h,
                                               __half* __restrict__ wf_l, int K, int C, int R, int S,
                                               size_t first, size_t step) {
  const size_t total = static_cast<size_t>(K) * C * R * S;
  for (size_t t = first; t < total; t += step) {
    const int c = static_cast<int>(t % C);
    size_t u = t / C;
    const int s = static_cast<int>(u % S); u /= S;
    const int r = static_cast<int>(u % R);
    const int k = static_cast<int>(u / R);
    const float v = w[((static_cast<size_t>(k) * C + c) * R + r) * S + s];
    const __half h = __float2half_rn(v);
    wf_h[t] = h;
    wf_l[t] = __float2half_rn(v - __half2float(h));  // hi + lo == v to ~2^-22
  }
}
__global__ void pack_fwd_kernel(const float* __restrict__ w, __half* __restrict__ wf_h,
                                __half* __restrict__ wf_l, int K, int C, int R, int S) {
  pack_fwd_range(w, wf_h, wf_l, K, C, R, S, static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x,
                 static_cast<size_t>(gridDim.x) * blockDim.x);
}
// data-gradient operand (flipped taps, in/out channels swapped):
//   wd[c][((R-1-r)*S + (S-1-s))*K + k] = w[k][c][r][s]
__device__ __forceinline__ void pack_dgrad_range(const float* __restrict__ w, float* __restrict__ wd,
                                                 int K, int C, int R, int S, size_t first, size_t step) {
  const size_t total = static_cast<size_t>(K) * C * R * S;
  for (size_t t = first; t < total; t += step) {
    const int k = static_cast<int>(t % K);
    size_t u = t / K;
    const int s2 = static_cast<int>(u % S); u /= S;
    const int r2 = static_cast<int>(u % R);
    const int c = static_cast<int>(u / R);
    wd[t] = tf32_rn(w[((static_cast<size_t>(k) * C + c) * R + (R - 1 - r2)) * S + (S - 1 - s2)]);
  }
}
__global__ void pack_dgrad_kernel(const float* __restrict__ w, float* __restrict__ wd, int K, int C,
                                  int R, int S) {
  pack_dgrad_range(w, wd, K, C, R, S, static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x,
                   static_cast<size_t>(gridDim.x) * blockDim.x);
}
// Stride-2 3x3 (pad 1) data gradient, split by output-pixel parity (ph, pw): only the taps
// with r = h + 1 (mod 2), s = w + 1 (mod 2) reach dx[h, w], so each class is a small stride-1
// convolution over dY with 1, 2, 2 or 4 taps:
//   dx[2i+ph, 2j+pw, c] = sum_{a < nr(ph), b < ns(pw), k} dY[i+a, j+b, k] * w[k][c][rr(ph,a)][ss(pw,b)]
//   ph = 0: one tap (a = 0 -> r = 1);  ph = 1: two taps (a = 0 -> r = 2, a = 1 -> r = 0)
// The four packs [C][(a*ns + b)*K + k] are stored back to back (9 * C * K floats in total, class
// order (0,0), (0,1), (1,0), (1,1)).
__global__ void pack_dgrad_s2_kernel(const float* __restrict__ w, float* __restrict__ wd, int K,
                                     int C) {
  const size_t per_tap = static_cast<size_t>(C) * K;
  const size_t total = 9 * per_tap;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // class offsets in taps: (0,0):0 [1 tap], (0,1):1 [2], (1,0):3 [2], (1,1):5 [4]
    const int g = static_cast<int>(t / per_tap);
    const int cls = g < 1 ? 0 : (g < 3 ? 1 : (g < 5 ? 2 : 3));
    const int first = cls == 0 ? 0 : (cls == 1 ? 1 : (cls == 2 ? 3 : 5));
    const int ntap = cls == 0 ? 1 : (cls == 3 ? 4 : 2);
    const int ph = cls >> 1, pw = cls & 1;
    const int ns = pw ? 2 : 1;
    const size_t u = t - first * per_tap;         // index inside the class pack [C][ntap*K]
    const int k = static_cast<int>(u % K);
    const int tap = static_cast<int>((u / K) % ntap);
    const int c = static_cast<int>(u / (static_cast<size_t>(K) * ntap));
    const int a = tap / ns, b = tap - a * ns;
    const int r = ph ? (a == 0 ? 2 : 0) : 1;
    const int s = pw ? (b == 0 ? 2 : 0) : 1;
    wd[t] = tf32_rn(w[((static_cast<size_t>(k) * C + c) * 3 + r) * 3 + s]);
  }
}
// The same nine (class, tap) blocks as ONE K-major matrix [C][9*K] (uniform row pitch): the B
// operand of the merged stride-2 data-gradient kernel (conv_igemm.cuh, S2M), whose K loop walks the
// blocks in this order.
__device__ __forceinline__ void pack_dgrad_s2m_range(const float* __restrict__ w, float* __restrict__ wd,
                                                     int K, int C, size_t first, size_t step) {
  const size_t total = static_cast<size_t>(9) * C * K;
  for (size_t t = first; t < total; t += step) {
    const int k = static_cast<int>(t % K);
    const int e = static_cast<int>((t / K) % 9);
    const int c = static_cast<int>(t / (static_cast<size_t>(9) * K));
    const int cls = e < 1 ? 0 : (e < 3 ? 1 : (e < 5 ? 2 : 3));
    const int tap = e - (cls == 0 ? 0 : (cls == 1 ? 1 : (cls == 2 ? 3 : 5)));
    const int ph = cls >> 1, pw = cls & 1;
    const int ns = pw ? 2 : 1;
    const int a = tap / ns, b = tap - a * ns;
    const int r = ph ? (a == 0 ? 2 : 0) : 1;
    const int s = pw ? (b == 0 ? 2 : 0) : 1;
    wd[t] = tf32_rn(w[((static_cast<size_t>(k) * C + c) * 3 + r) * 3 + s]);
  }
}
__global__ void pack_dgrad_s2m_kernel(const float* __restrict__ w, float* __restrict__ wd, int K,
                                      int C) {
  pack_dgrad_s2m_range(w, wd, K, C, static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x,
                       static_cast<size_t>(gridDim.x) * blockDim.x);
}

// Many packs in one launch (the ~40 weight packs a training step needs are ~2-8 us kernels each:
// launch gaps, not work).  Job j owns blocks [block0[j], block0[j+1]) of the grid.
constexpr int kMaxPackJobs = 64;
struct PackJob {
  const float* w;
  void* d0;
  void* d1;
  int K, C, R, S;
  int kind;    // 0: forward (hi, lo) FP16 pair, 1: data gradient, 2: merged stride-2 data gradient
  int block0;  // first block of the job
};
struct PackTable {
  PackJob job[kMaxPackJobs];
  int n;
  int total_blocks;
};
__global__ void __launch_bounds__(256) pack_multi_kernel(const __grid_constant__ PackTable tbl) {
  int j = 0;
  while (j + 1 < tbl.n && static_cast<int>(blockIdx.x) >= tbl.job[j + 1].block0) ++j;
  const PackJob& q = tbl.job[j];
  const int nblocks = (j + 1 < tbl.n ? tbl.job[j + 1].block0 : tbl.total_blocks) - q.block0;
  const size_t first = static_cast<size_t>(blockIdx.x - q.block0) * 256 + threadIdx.x;
  const size_t step = static_cast<size_t>(nblocks) * 256;
  if (q.kind == 0)
    pack_fwd_range(q.w, static_cast<__half*>(q.d0), static_cast<__half*>(q.d1), q.K, q.C, q.R, q.S, first, step);
  else if (q.kind == 1)
    pack_dgrad_range(q.w, static_cast<float*>(q.d0), q.K, q.C, q.R, q.S, first, step);
  else
    pack_dgrad_s2m_range(q.w, static_cast<float*>(q.d0), q.K, q.C, first, step);
}
// weight-gradient result back to the parameter layout: dw[k][c][r][s] = dwf[k][(r*S+s)*C + c]
// (accumulate: += into a gradient slot several passes over the same weights feed)
// planes > 1: dwf holds that many [K][R*S*C] partial planes (deterministic split-K), summed here
// in plane order
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwf, float* __restrict__ dw, int K,
                                    int C, int R, int S, int accumulate, int planes) {
  const size_t total = static_cast<size_t>(K) * C * R * S;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int s = static_cast<int>(t % S);
    size_t u = t / S;
    const int r = static_cast<int>(u % R); u /= R;
    const int c = static_cast<int>(u % C);
    const int k = static_cast<int>(u / C);
    const size_t src = (static_cast<size_t>(k) * R * S + r * S + s) * C + c;
    float v = dwf[src];
    for (int pl = 1; pl < planes; ++pl) v += dwf[static_cast<size_t>(pl) * total + src];
    dw[t] = accumulate ? dw[t] + v : v;
  }
}

static unsigned pack_grid(size_t total) {
  size_t b = (total + 255) / 256;
  if (b > 2048) b = 2048;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}
#define B2N_PACK_LAUNCH(NAME, KERNEL)                                                       \
  int NAME(const float* src, float* dst, int K, int C, int R, int S, cudaStream_t stream) { \
    const size_t total = static_cast<size_t>(K) * C * R * S;                               \
    KERNEL<<<pack_grid(total), 256, 0, stream>>>(src, dst, K, C, R, S);                     \
    cudaError_t e = cudaGetLastError();                                                     \
    if (e != cudaSuccess) return set_error(#NAME ": %s", cudaGetErrorString(e));            \
    return 0;                                                                               \
  }
int launch_pack_fwd(const float* src, __half* dst_h, __half* dst_l, int K, int C, int R, int S,
                    cudaStream_t stream) {
  const size_t total = static_cast<size_t>(K) * C * R * S;
  pack_fwd_kernel<<<pack_grid(total), 256, 0, stream>>>(src, dst_h, dst_l, K, C, R, S);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("launch_pack_fwd: %s", cudaGetErrorString(e));
  return 0;
}
B2N_PACK_LAUNCH(launch_pack_dgrad, pack_dgrad_kernel)
#undef B2N_PACK_LAUNCH
int launch_unpack_wgrad(const float* src, float* dst, int K, int C, int R, int S, int accumulate,
                        int planes, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(K) * C * R * S;
  if (planes < 1) return set_error("unpack_wgrad: planes must be >= 1");
  unpack_wgrad_kernel<<<pack_grid(total), 256, 0, stream>>>(src, dst, K, C, R, S, accumulate, planes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("launch_unpack_wgrad: %s", cudaGetErrorString(e));
  return 0;
}

int launch_pack_multi(const float* const* w, void* const* d0, void* const* d1, const int* kind,
                      const int* K, const int* C, const int* R, const int* S, int n, cudaStream_t stream) {
  for (int base = 0; base < n; base += kMaxPackJobs) {
    PackTable tbl;
    tbl.n = n - base < kMaxPackJobs ? n - base : kMaxPackJobs;
    int blocks = 0;
    for (int i = 0; i < tbl.n; ++i) {
      const int g = base + i;
      PackJob& q = tbl.job[i];
      if (w[g] == nullptr || d0[g] == nullptr || kind[g] < 0 || kind[g] > 2 || (kind[g] == 0 && d1[g] == nullptr))
        return set_error("pack_multi: bad job %d", g);
      if (kind[g] == 2 && (R[g] != 3 || S[g] != 3)) return set_error("pack_multi: job %d: the stride-2 pack is 3x3", g);
      q.w = w[g]; q.d0 = d0[g]; q.d1 = d1[g];
      q.K = K[g]; q.C = C[g]; q.R = R[g]; q.S = S[g];
      q.kind = kind[g];
      q.block0 = blocks;
      const size_t total = static_cast<size_t>(K[g]) * C[g] * R[g] * S[g];
      size_t b = (total + 2047) / 2048;   // eight elements per thread
      if (b > 256) b = 256;
      if (b < 1) b = 1;
      blocks += static_cast<int>(b);
    }
    tbl.total_blocks = blocks;
    pack_multi_kernel<<<blocks, 256, 0, stream>>>(tbl);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("launch_pack_multi: %s", cudaGetErrorString(e));
  }
  return 0;
}

int launch_pack_dgrad_s2m(const float* src, float* dst, int K, int C, cudaStream_t stream) {
  pack_dgrad_s2m_kernel<<<pack_grid(static_cast<size_t>(9) * C * K), 256, 0, stream>>>(src, dst, K, C);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("launch_pack_dgrad_s2m: %s", cudaGetErrorString(e));
  return 0;
}
int launch_pack_dgrad_s2(const float* src, float* dst, int K, int C, cudaStream_t stream) {
  pack_dgrad_s2_kernel<<<pack_grid(static_cast<size_t>(9) * C * K), 256, 0, stream>>>(src, dst, K, C);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("launch_pack_dgrad_s2: %s", cudaGetErrorString(e));
  return 0;
}


}  // namespace b2n
