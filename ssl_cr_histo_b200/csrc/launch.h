// Internal host-side launch interfaces shared by the C-ABI layer (api.cu) and the
// stand-alone device tests (devtest.cu).  Not part of the public ABI (see include/b2n.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2n {

// printf-style error recorder; always returns a non-zero code so callers can `return set_error(..)`.
int set_error(const char* fmt, ...);
const char* last_error();
int device_sm_count();

// Forward-style convolution: out[n,p,q,:] = sum_taps x[n, p*stride - pad_lo + r, ...] * w.
// x is NHWC fp32 [N,H,W,Cin]; w is packed K-major [Cout][R*S*Cin]; out is NHWC [N,P,Q,Cout].
struct ConvArgs {
  const float* x = nullptr;
  const float* w = nullptr;
  float* out = nullptr;
  int N = 0, H = 0, W = 0, Cin = 0, Cout = 0, R = 0, S = 0, stride = 1;
  int pad_h_lo = 0, pad_h_hi = 0, pad_w_lo = 0, pad_w_hi = 0;
  const float* scale = nullptr;
  const float* shift = nullptr;
  const float* resid = nullptr;
  const float* mask = nullptr;
  int relu = 0;
  int round_tf32 = 0;
  double* stats = nullptr;
  int force_block_n = 0;  // 0 = heuristic
};
int launch_conv(const ConvArgs& a, cudaStream_t stream);

// Weight gradient: dw[k][(r*S+s)*Cin + c] += sum_pixels dy[pix][k] * x[pix shifted by tap][c].
// dw must be zero-initialised by the caller (split-K partial sums are accumulated atomically).
struct WgradArgs {
  const float* x = nullptr;   // NHWC [N,H,W,Cin]
  const float* dy = nullptr;  // NHWC [N,P,Q,Cout]
  float* dw = nullptr;        // [Cout][R*S*Cin]
  int N = 0, H = 0, W = 0, Cin = 0, Cout = 0, R = 0, S = 0, stride = 1;
  int pad_h_lo = 0, pad_h_hi = 0, pad_w_lo = 0, pad_w_hi = 0;
  int force_splits = 0;  // 0 = heuristic
};
int launch_wgrad(const WgradArgs& a, cudaStream_t stream);

}  // namespace b2n
