// Weight re-layout kernels.  Parameters stay in PyTorch's (K,C,R,S) layout so reference
// checkpoints load unchanged; the tensor-core kernels consume K-major
// packs with the reduction index (tap, channel) contiguous, values rounded to TF32.
#include <cuda_fp16.h>

#include "launch.h"
#include "ptx.cuh"

namespace b2n {

// forward operand (FP16 hi/lo pair):  wf[k][(r*S+s)*C + c] = w[k][c][r][s]
__global__ void pack_fwd_kernel(const float* __restrict__ w, __half* __restrict__ wf_h,
                                __half* __restrict__ wf_l, int K, int C, int R, int S) {
  const size_t total = static_cast<size_t>(K) * C * R * S;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
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
// data-gradient operand (flipped taps, in/out channels swapped):
//   wd[c][((R-1-r)*S + (S-1-s))*K + k] = w[k][c][r][s]
__global__ void pack_dgrad_kernel(const float* __restrict__ w, float* __restrict__ wd, int K, int C,
                                  int R, int S) {
  const size_t total = static_cast<size_t>(K) * C * R * S;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(t % K);
    size_t u = t / K;
    const int s2 = static_cast<int>(u % S); u /= S;
    const int r2 = static_cast<int>(u % R);
    const int c = static_cast<int>(u / R);
    wd[t] = tf32_rn(w[((static_cast<size_t>(k) * C + c) * R + (R - 1 - r2)) * S + (S - 1 - s2)]);
  }
}
// weight-gradient result back to the parameter layout: dw[k][c][r][s] = dwf[k][(r*S+s)*C + c]
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwf, float* __restrict__ dw, int K,
                                    int C, int R, int S) {
  const size_t total = static_cast<size_t>(K) * C * R * S;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int s = static_cast<int>(t % S);
    size_t u = t / S;
    const int r = static_cast<int>(u % R); u /= R;
    const int c = static_cast<int>(u % C);
    const int k = static_cast<int>(u / C);
    dw[t] = dwf[(static_cast<size_t>(k) * R * S + r * S + s) * C + c];
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
B2N_PACK_LAUNCH(launch_unpack_wgrad, unpack_wgrad_kernel)
#undef B2N_PACK_LAUNCH

}  // namespace b2n
