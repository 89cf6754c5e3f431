// Internal host-side launch interfaces shared by the C-ABI layer (api.cu) and the
// stand-alone device tests (devtest.cu).  Not part of the public ABI (see include/b2n.h).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2n {

// printf-style error recorder; always returns a non-zero code so callers can `return set_error(..)`.
int set_error(const char* fmt, ...);
const char* last_error();
int device_sm_count();   // of the current device (cached per device)
int elementwise_blocks_per_sm();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device, and the entry points may be
// called from several host threads (one per GPU under the reference's nn.DataParallel): every
// launcher keeps one PerDeviceMax per kernel instantiation -- the largest value already set on each
// device (lock-free; setting the attribute twice is harmless, so a lost race only repeats the call).
constexpr int kMaxDevices = 64;
struct PerDeviceMax {
  int v[kMaxDevices] = {};
  // true when `want` exceeds what has been configured on the current device (then call set())
  bool needs(int want, int* dev_out) const {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev;
    if (dev < 0 || dev >= kMaxDevices) return true;
    return __atomic_load_n(&v[dev], __ATOMIC_ACQUIRE) < want;
  }
  void set(int dev, int want) {
    if (dev < 0 || dev >= kMaxDevices) return;
    int cur = __atomic_load_n(&v[dev], __ATOMIC_RELAXED);
    while (cur < want && !__atomic_compare_exchange_n(&v[dev], &cur, want, false, __ATOMIC_RELEASE,
                                                       __ATOMIC_RELAXED)) {
    }
  }
};

// Forward-style convolution: out[n,p,q,:] = sum_taps x[n, p*stride - pad_lo + r, ...] * w.
// x is NHWC fp32 [N,H,W,Cin]; w is packed K-major [Cout][R*S*Cin]; out is NHWC [N,P,Q,Cout].
struct ConvArgs {
  const float* x = nullptr;      // plain mode: TF32 operands in fp32 containers
  const float* w = nullptr;
  const float* x2 = nullptr;     // merged stride-2 data gradient only: dY of the block's 1x1 shortcut
  const float* w2 = nullptr;     // conv and its data-gradient pack [C][K] (accumulated into class (0,0))
  const __half* x_h = nullptr;   // split mode: error-compensated (hi, lo) FP16 operand pairs
  const __half* x_l = nullptr;
  const __half* w_h = nullptr;
  const __half* w_l = nullptr;
  float* out = nullptr;          // fp32 result and / or ...
  __half* out_h = nullptr;       // ... the result as a (hi, lo) FP16 pair
  __half* out_l = nullptr;
  int N = 0, H = 0, W = 0, Cin = 0, Cout = 0, R = 0, S = 0, stride = 1;
  int pad_h_lo = 0, pad_h_hi = 0, pad_w_lo = 0, pad_w_hi = 0;
  const float* scale = nullptr;
  const float* shift = nullptr;
  const float* resid = nullptr;
  const __half* resid_h = nullptr;
  const __half* resid_l = nullptr;
  const float* mask = nullptr;
  const float* gate = nullptr;        // result zeroed where gate <= 0 (addressed like the output)
  const float* bnb_y = nullptr;       // BatchNorm-backward sums of the result into `stats` ...
  const float* bnb_mean = nullptr;
  const float* bnb_invstd = nullptr;
  const float* bnb_scale = nullptr;   // ... after the ReLU gate fmaf(y, scale, shift) > 0 (optional)
  const float* bnb_shift = nullptr;
  int relu = 0;
  int round_tf32 = 0;
  double* stats = nullptr;
  int force_block_n = 0;  // 0 = heuristic
  int o_step = 0, o_h0 = 0, o_w0 = 0, o_H = 0, o_W = 0;  // strided output placement (0 = dense)
  int a_tiled2d = 0;                  // experiment: 1x1/s1 conv with A loaded in tiled mode
  int no_resident_weights = 0;        // force the streamed-weights variant (tests)
  int no_halo = 0;                    // force one box per filter tap (tests)
  const int* a_lo_nonzero = nullptr;  // split mode: device flag, 0 => x_l is all zero (skipped)
};
int launch_conv(const ConvArgs& a, cudaStream_t stream);
// Merged stride-2 3x3/pad-1 data gradient (all four output-parity classes in one launch):
// x = dY [N,H,W,Cin] (H, W: its spatial size; Cin: its channels = the forward conv's Cout),
// w = b2n_pack_weight_dgrad_s2m pack [Cout][9*Cin], out = dX [N,o_H,o_W,Cout]; resid (optional) is
// added on the even-even pixels only (the 1x1 shortcut conv's gradient, already in `out`), gate
// (optional) zeroes dX where it is <= 0.
int launch_conv_dgrad_s2(const ConvArgs& a, cudaStream_t stream);

// Weight gradient: dw[k][(r*S+s)*Cin + c] += sum_pixels dy[pix][k] * x[pix shifted by tap][c].
// dw must be zero-initialised by the caller (split-K partial sums are accumulated atomically).
struct WgradArgs {
  const float* x = nullptr;   // NHWC [N,H,W,Cin]
  const float* dy = nullptr;  // NHWC [N,P,Q,Cout]
  float* dw = nullptr;        // [Cout][R*S*Cin]
  int N = 0, H = 0, W = 0, Cin = 0, Cout = 0, R = 0, S = 0, stride = 1;
  int pad_h_lo = 0, pad_h_hi = 0, pad_w_lo = 0, pad_w_hi = 0;
  int force_splits = 0;  // 0 = heuristic
  int no_halo = 0;       // force the one-box-per-tap kernel (tests)
  int deterministic = 0; // dw holds wgrad_planes() partial planes, stored without atomics
  int x_channels = 0;    // channels x stores per pixel (0 = Cin); fewer than Cin: the missing ones are
                         // zero-filled by the TMA unit (the stem's 12 real of 32 reduction channels)
};
int launch_wgrad(const WgradArgs& a, cudaStream_t stream);
// number of [Cout][R*S*Cin] planes a deterministic launch of this shape writes on the current
// device (>= 1), or -1 with the error set
int wgrad_planes(const WgradArgs& a);


// ---- element-wise / small kernels (bn.cu, stem_pool.cu, pack.cu, linear.cu, loss_lerp.cu) ----
int launch_bn_finalize(const double* stats, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float* scale, float* shift,
                       float* mean, float* invstd, float* inv_gamma, int C, double count,
                       float momentum, float eps, int n_updates, cudaStream_t stream);
int launch_bn_fold_eval_multi(const float* const* gamma, const float* const* beta, const float* const* rm,
                              const float* const* rv, float* const* scale, float* const* shift,
                              const int* C, const float* eps, int n, cudaStream_t stream);
int launch_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                        float* scale, float* shift, int C, float eps, cudaStream_t stream);
int launch_bn_apply(const float* y, const float* scale, const float* shift, const float* res32,
                    const float* res_scale, const float* res_shift, const __half* res_h,
                    const __half* res_l, float* out32, __half* out_h, __half* out_l,
                    long long rows, int C, int relu, int round_tf32, cudaStream_t stream);
int launch_bn_bwd_reduce(const float* g, const float* mask, const float* y, const float* mean,
                         const float* invstd, const float* gate_scale, const float* gate_shift,
                         double* sums, long long rows, int C, cudaStream_t stream);
int launch_bn_bwd_apply(const float* g, const float* mask, const float* y, const float* mean,
                        const float* invstd, const float* gamma, const float* gate_scale,
                        const float* gate_shift, const double* sums, float* dy, float* dgamma,
                        float* dbeta, long long rows, int C, int round_tf32, int accumulate,
                        cudaStream_t stream);
int launch_upsample_zero(const float* dy, float* up, int N, int P, int Q, int H, int W, int C,
                         cudaStream_t stream);
int launch_stem_pack_input(const float* x, __half* xs_h, __half* xs_l, float* xs32,
                           int* lo_nonzero, int N, int H, int W, cudaStream_t stream);
int launch_stem_pack_input_u8(const unsigned char* x, __half* xs_h, float* xs32, int N, int H, int W,
                              cudaStream_t stream);
int launch_stem_pack_weight(const float* w, __half* ws_h, __half* ws_l, int K,
                            cudaStream_t stream);
int launch_stem_unpack_wgrad(const float* dws, float* dw, int K, int accumulate, int planes,
                             cudaStream_t stream);
int launch_bn_relu_maxpool(const float* y, const float* scale, const float* shift, float* a32,
                           __half* a_h, __half* a_l, unsigned char* idx, int N, int H, int W,
                           int C, cudaStream_t stream);
int launch_maxpool_relu_bwd(const float* ga, const unsigned char* idx, const float* y,
                            const float* scale, const float* shift, float* gz, int N, int H, int W,
                            int C, cudaStream_t stream);
int launch_pool_bn_bwd_reduce(const float* ga, const unsigned char* idx, const float* y,
                              const float* scale, const float* shift, const float* mean,
                              const float* invstd, double* sums, int N, int H, int W, int C,
                              cudaStream_t stream);
int launch_pool_bn_bwd_apply(const float* ga, const unsigned char* idx, const float* y,
                             const float* scale, const float* shift, const float* mean,
                             const float* invstd, const float* gamma, const double* sums, float* dy,
                             float* dgamma, float* dbeta, int N, int H, int W, int C, int round_tf32,
                             int accumulate, cudaStream_t stream);
int launch_avgpool_fwd(const __half* a_h, const __half* a_l, float* e, int N, int HW, int C,
                       cudaStream_t stream);
int launch_avgpool_bwd(const float* ge, const float* gate, float* g, int N, int HW, int C,
                       cudaStream_t stream);
int launch_pack_multi(const float* const* w, void* const* d0, void* const* d1, const int* kind,
                      const int* K, const int* C, const int* R, const int* S, int n, cudaStream_t stream);
int launch_pack_fwd(const float* src, __half* dst_h, __half* dst_l, int K, int C, int R, int S,
                    cudaStream_t stream);
int launch_pack_dgrad(const float* src, float* dst, int K, int C, int R, int S,
                      cudaStream_t stream);
int launch_pack_dgrad_s2(const float* src, float* dst, int K, int C, cudaStream_t stream);
int launch_pack_dgrad_s2m(const float* src, float* dst, int K, int C, cudaStream_t stream);
int launch_unpack_wgrad(const float* src, float* dst, int K, int C, int R, int S, int accumulate,
                        int planes, cudaStream_t stream);
int launch_linear_fwd(const float* x, long long ldx, const float* w, long long ldw, const float* b,
                      float* y, long long ldy, int rows, int in_f, int out_f, int relu,
                      int accumulate, cudaStream_t stream);
int launch_linear_bwd_data(const float* dy, long long lddy, const float* w, long long ldw,
                           float* dx, long long lddx, const float* mask, int rows, int in_f,
                           int out_f, int accumulate, cudaStream_t stream);
int launch_linear_bwd_weight(const float* dy, long long lddy, const float* x, long long ldx,
                             float* dw, long long lddw, int rows, int in_f, int out_f,
                             int accumulate, cudaStream_t stream);
int launch_cols_replicate(float* y, long long ld, int rows, int width, int copies, cudaStream_t stream);
int launch_cols_sum(const float* dy, long long ld, float* out, int rows, int width, int copies,
                    cudaStream_t stream);
int launch_colsum(const float* dy, long long lddy, float* db, int rows, int out_f, int accumulate,
                  cudaStream_t stream);
int launch_fused_loss(int mode, const float* logits_x, const long long* targets_i,
                      const float* targets_f, const float* logits_u_w, const float* logits_u_s,
                      int rows_x, int rows_u, int C, float lambda_u, float* losses,
                      float* dlogits_x, float* dlogits_u, long long* argmax_x,
                      long long* pseudo_out, cudaStream_t stream);
int launch_softmax_last(const float* logits, float* out, int rows, int C, cudaStream_t stream);
int launch_lerp_multi(float* const* dst, float* const* src, const long long* numel, int n,
                      float alpha, int write_back, cudaStream_t stream);

// ---- GPU augmentation of uint8 (N,3,H,W) batches (augment.cu) ----
int launch_aug_flip_crop(const unsigned char* src, unsigned char* dst, const int* top, const int* left,
                         const int* flip, int N, int Hs, int Ws, int H, int W, cudaStream_t stream);
int launch_aug_brightness_contrast(const unsigned char* src, unsigned char* dst, const float* alpha,
                                   const float* offset, const int* apply, int N, int H, int W,
                                   cudaStream_t stream);
int launch_aug_image_mean(const unsigned char* src, float* mean, int N, int H, int W, cudaStream_t stream);
int launch_aug_hsv_shift(const unsigned char* src, unsigned char* dst, const int* dh, const int* ds,
                         const int* dv, const int* apply, int N, int H, int W, cudaStream_t stream);
int launch_aug_add_noise(const unsigned char* src, unsigned char* dst, const float* noise, const int* apply,
                         int N, int H, int W, cudaStream_t stream);
int launch_aug_box_blur(const unsigned char* src, unsigned char* dst, const int* ksize, const int* apply,
                        int N, int H, int W, cudaStream_t stream);
int launch_aug_hed_jitter(const unsigned char* src, unsigned char* dst, const float* delta, const int* apply,
                          int N, int H, int W, cudaStream_t stream);
int launch_aug_warp_affine(const unsigned char* src, unsigned char* dst, const float* minv, const int* apply,
                           int N, int Hs, int Ws, int H, int W, int clamp_border, cudaStream_t stream);

int launch_adam_multi(float* const* p, const float* const* g, float* const* exp_avg,
                      float* const* exp_avg_sq, const long long* numel, int n, double lr, double beta1,
                      double beta2, double eps, double weight_decay, long long step,
                      long long* step_dev, double grad_scale, cudaStream_t stream);
int launch_sgd_multi(float* const* p, const float* const* g, float* const* momentum_buf,
                     const long long* numel, int n, double lr, double momentum, double weight_decay,
                     int nesterov, int first_step, double grad_scale, cudaStream_t stream);

}  // namespace b2n
