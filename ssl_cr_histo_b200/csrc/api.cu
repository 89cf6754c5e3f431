// extern "C" surface of libb2n.so -- see include/b2n.h for the contract of every entry point.
#include "../../include/b2n.h"

#include <cuda_fp16.h>

#include "launch.h"

using namespace b2n;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline __half* H16(b2n_half* p) { return reinterpret_cast<__half*>(p); }
static inline const __half* H16(const b2n_half* p) { return reinterpret_cast<const __half*>(p); }

// number of kernels enqueued through this library since load (bench.py's gpu_launches)
static unsigned long long g_launches = 0;
static inline int counted(int rc, int kernels = 1) {
  if (rc == 0) __atomic_fetch_add(&g_launches, (unsigned long long)kernels, __ATOMIC_RELAXED);
  return rc;
}

extern "C" {

int b2n_version(void) { return B2N_ABI_VERSION; }
unsigned long long b2n_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
const char* b2n_last_error(void) { return last_error(); }

int b2n_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return major == 10 ? 1 : 0;
}

int b2n_conv_fwd(const float* x, const b2n_half* x_h, const b2n_half* x_l, const float* w_packed,
                 const b2n_half* w_h, const b2n_half* w_l, float* y, b2n_half* y_h, b2n_half* y_l,
                 int N, int H, int W, int Cin, int Cout, int R, int Sf, int stride, int pad_h_lo,
                 int pad_h_hi, int pad_w_lo, int pad_w_hi, const float* scale, const float* shift,
                 const float* resid, const b2n_half* resid_h, const b2n_half* resid_l,
                 const float* mask, int relu, int round_tf32, double* stats,
                 const int* x_l_nonzero, int o_step, int o_h0, int o_w0, int o_H, int o_W,
                 const float* gate, const float* bnb_y, const float* bnb_mean,
                 const float* bnb_invstd, const float* bnb_scale, const float* bnb_shift,
                 void* stream) {
  ConvArgs a;
  a.gate = gate; a.bnb_y = bnb_y; a.bnb_mean = bnb_mean; a.bnb_invstd = bnb_invstd;
  a.bnb_scale = bnb_scale; a.bnb_shift = bnb_shift;
  a.a_lo_nonzero = x_l_nonzero;
  a.o_step = o_step; a.o_h0 = o_h0; a.o_w0 = o_w0; a.o_H = o_H; a.o_W = o_W;
  a.x = x; a.w = w_packed;
  a.x_h = H16(x_h); a.x_l = H16(x_l); a.w_h = H16(w_h); a.w_l = H16(w_l);
  a.out = y; a.out_h = H16(y_h); a.out_l = H16(y_l);
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.R = R; a.S = Sf; a.stride = stride;
  a.pad_h_lo = pad_h_lo; a.pad_h_hi = pad_h_hi; a.pad_w_lo = pad_w_lo; a.pad_w_hi = pad_w_hi;
  a.scale = scale; a.shift = shift; a.resid = resid; a.resid_h = H16(resid_h);
  a.resid_l = H16(resid_l); a.mask = mask;
  a.relu = relu; a.round_tf32 = round_tf32; a.stats = stats;
  return counted(launch_conv(a, S(stream)));
}

int b2n_conv_wgrad(const float* x, const float* dy, float* dw_packed, int N, int H, int W, int Cin,
                   int Cout, int R, int Sf, int stride, int pad_h_lo, int pad_h_hi, int pad_w_lo,
                   int pad_w_hi, int deterministic, int x_channels, void* stream) {
  if (!x || !dy || !dw_packed) return set_error("b2n_conv_wgrad: null tensor");
  WgradArgs a;
  a.deterministic = deterministic;
  a.x_channels = x_channels;
  a.x = x; a.dy = dy; a.dw = dw_packed;
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.R = R; a.S = Sf; a.stride = stride;
  a.pad_h_lo = pad_h_lo; a.pad_h_hi = pad_h_hi; a.pad_w_lo = pad_w_lo; a.pad_w_hi = pad_w_hi;
  return counted(launch_wgrad(a, S(stream)));
}

int b2n_conv_dgrad_s2(const float* dy, const float* w_packed, float* dx, int N, int P, int Q, int K,
                      int C, int H, int W, const float* resid, const float* gate, void* stream) {
  ConvArgs a;
  a.x = dy; a.w = w_packed; a.out = dx;
  a.N = N; a.H = P; a.W = Q; a.Cin = K; a.Cout = C; a.o_H = H; a.o_W = W;
  a.resid = resid; a.gate = gate;
  return counted(launch_conv_dgrad_s2(a, S(stream)));
}
int b2n_conv_dgrad_s2_sc(const float* dy, const float* w_packed, const float* dy_sc, const float* w_sc_packed,
                         float* dx, int N, int P, int Q, int K, int C, int H, int W, const float* gate,
                         void* stream) {
  if (!dy_sc || !w_sc_packed) return set_error("b2n_conv_dgrad_s2_sc: null shortcut operand");
  ConvArgs a;
  a.x = dy; a.w = w_packed; a.x2 = dy_sc; a.w2 = w_sc_packed; a.out = dx;
  a.N = N; a.H = P; a.W = Q; a.Cin = K; a.Cout = C; a.o_H = H; a.o_W = W;
  a.gate = gate;
  return counted(launch_conv_dgrad_s2(a, S(stream)));
}
int b2n_pack_weight_dgrad_s2m(const float* w, float* wp, int K, int C, void* stream) {
  return counted(launch_pack_dgrad_s2m(w, wp, K, C, S(stream)));
}

int b2n_conv_wgrad_planes(int N, int H, int W, int Cin, int Cout, int R, int Sf, int stride,
                          int pad_h_lo, int pad_h_hi, int pad_w_lo, int pad_w_hi) {
  WgradArgs a;
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.R = R; a.S = Sf; a.stride = stride;
  a.pad_h_lo = pad_h_lo; a.pad_h_hi = pad_h_hi; a.pad_w_lo = pad_w_lo; a.pad_w_hi = pad_w_hi;
  return wgrad_planes(a);
}

int b2n_pack_weight_fwd(const float* w, b2n_half* wp_h, b2n_half* wp_l, int K, int C, int R, int Sf,
                        void* stream) {
  return counted(launch_pack_fwd(w, H16(wp_h), H16(wp_l), K, C, R, Sf, S(stream)));
}
int b2n_pack_weights_multi(const float* const* w, void* const* dst0, void* const* dst1, const int* kind,
                           const int* K, const int* C, const int* R, const int* Sf, int n, void* stream) {
  if (n < 0 || (n > 0 && (!w || !dst0 || !dst1 || !kind || !K || !C || !R || !Sf)))
    return set_error("b2n_pack_weights_multi: bad args");
  if (n == 0) return 0;
  return counted(launch_pack_multi(w, dst0, dst1, kind, K, C, R, Sf, n, S(stream)), (n + 63) / 64);
}
int b2n_pack_weight_dgrad(const float* w, float* wp, int K, int C, int R, int Sf, void* stream) {
  return counted(launch_pack_dgrad(w, wp, K, C, R, Sf, S(stream)));
}
int b2n_pack_weight_dgrad_s2(const float* w, float* wp, int K, int C, void* stream) {
  return counted(launch_pack_dgrad_s2(w, wp, K, C, S(stream)));
}
int b2n_unpack_wgrad(const float* dwp, float* dw, int K, int C, int R, int Sf, int accumulate,
                     int planes, void* stream) {
  return counted(launch_unpack_wgrad(dwp, dw, K, C, R, Sf, accumulate, planes, S(stream)));
}

int b2n_stem_pack_input(const float* x, b2n_half* xs_h, b2n_half* xs_l, float* xs32,
                        int* xs_l_nonzero, int N, int H, int W, void* stream) {
  return counted(launch_stem_pack_input(x, H16(xs_h), H16(xs_l), xs32, xs_l_nonzero, N, H, W,
                                        S(stream)));
}
int b2n_stem_pack_input_u8(const unsigned char* x, b2n_half* xs_h, float* xs32, int N, int H, int W,
                           void* stream) {
  return counted(launch_stem_pack_input_u8(x, H16(xs_h), xs32, N, H, W, S(stream)));
}
int b2n_stem_pack_weight(const float* w, b2n_half* ws_h, b2n_half* ws_l, int K, void* stream) {
  return counted(launch_stem_pack_weight(w, H16(ws_h), H16(ws_l), K, S(stream)));
}
int b2n_stem_unpack_wgrad(const float* dws, float* dw, int K, int accumulate, int planes,
                          void* stream) {
  return counted(launch_stem_unpack_wgrad(dws, dw, K, accumulate, planes, S(stream)));
}

int b2n_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float* scale, float* shift, float* mean, float* invstd,
                    float* inv_gamma, int C, double count, float momentum, float eps, int n_updates,
                    void* stream) {
  return counted(launch_bn_finalize(stats, gamma, beta, running_mean, running_var, scale, shift, mean,
                            invstd, inv_gamma, C, count, momentum, eps, n_updates, S(stream)));
}
int b2n_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                     float* scale, float* shift, int C, float eps, void* stream) {
  return counted(launch_bn_fold_eval(gamma, beta, rm, rv, scale, shift, C, eps, S(stream)));
}
int b2n_bn_fold_eval_multi(const float* const* gamma, const float* const* beta, const float* const* rm,
                           const float* const* rv, float* const* scale, float* const* shift, const int* C,
                           const float* eps, int n, void* stream) {
  if (n < 0 || (n > 0 && (!gamma || !beta || !rm || !rv || !scale || !shift || !C || !eps)))
    return set_error("b2n_bn_fold_eval_multi: bad args");
  if (n == 0) return 0;
  return counted(launch_bn_fold_eval_multi(gamma, beta, rm, rv, scale, shift, C, eps, n, S(stream)),
                 (n + 31) / 32);
}
int b2n_bn_apply(const float* y, const float* scale, const float* shift, const float* res32,
                 const float* res_scale, const float* res_shift, const b2n_half* res_h,
                 const b2n_half* res_l, float* out32, b2n_half* out_h, b2n_half* out_l,
                 long long rows, int C, int relu, int round_tf32, void* stream) {
  return counted(launch_bn_apply(y, scale, shift, res32, res_scale, res_shift, H16(res_h),
                                 H16(res_l), out32, H16(out_h), H16(out_l), rows, C, relu,
                                 round_tf32, S(stream)));
}
int b2n_bn_bwd_reduce(const float* g, const float* mask, const float* y, const float* mean,
                      const float* invstd, const float* gate_scale, const float* gate_shift,
                      double* sums, long long rows, int C, void* stream) {
  return counted(launch_bn_bwd_reduce(g, mask, y, mean, invstd, gate_scale, gate_shift, sums, rows, C,
                                      S(stream)));
}
int b2n_bn_bwd_apply(const float* g, const float* mask, const float* y, const float* mean,
                     const float* invstd, const float* gamma, const float* gate_scale,
                     const float* gate_shift, const double* sums, float* dy, float* dgamma,
                     float* dbeta, long long rows, int C, int round_tf32, int accumulate,
                     void* stream) {
  return counted(launch_bn_bwd_apply(g, mask, y, mean, invstd, gamma, gate_scale, gate_shift, sums, dy,
                                     dgamma, dbeta, rows, C, round_tf32, accumulate, S(stream)));
}
int b2n_upsample_zero(const float* dy, float* up, int N, int P, int Q, int H, int W, int C,
                      void* stream) {
  return counted(launch_upsample_zero(dy, up, N, P, Q, H, W, C, S(stream)));
}

int b2n_bn_relu_maxpool(const float* y, const float* scale, const float* shift, float* a32,
                        b2n_half* a_h, b2n_half* a_l, unsigned char* idx, int N, int H, int W,
                        int C, void* stream) {
  return counted(launch_bn_relu_maxpool(y, scale, shift, a32, H16(a_h), H16(a_l), idx, N, H, W, C,
                                        S(stream)));
}
int b2n_maxpool_relu_bwd(const float* ga, const unsigned char* idx, const float* y,
                         const float* scale, const float* shift, float* gz, int N, int H, int W,
                         int C, void* stream) {
  return counted(launch_maxpool_relu_bwd(ga, idx, y, scale, shift, gz, N, H, W, C, S(stream)));
}
int b2n_pool_bn_bwd_reduce(const float* ga, const unsigned char* idx, const float* y,
                           const float* scale, const float* shift, const float* mean,
                           const float* invstd, double* sums, int N, int H, int W, int C,
                           void* stream) {
  return counted(launch_pool_bn_bwd_reduce(ga, idx, y, scale, shift, mean, invstd, sums, N, H, W, C,
                                           S(stream)));
}
int b2n_pool_bn_bwd_apply(const float* ga, const unsigned char* idx, const float* y,
                          const float* scale, const float* shift, const float* mean,
                          const float* invstd, const float* gamma, const double* sums, float* dy,
                          float* dgamma, float* dbeta, int N, int H, int W, int C, int round_tf32,
                          int accumulate, void* stream) {
  return counted(launch_pool_bn_bwd_apply(ga, idx, y, scale, shift, mean, invstd, gamma, sums, dy,
                                          dgamma, dbeta, N, H, W, C, round_tf32, accumulate,
                                          S(stream)));
}
int b2n_avgpool_fwd(const b2n_half* a_h, const b2n_half* a_l, float* e, int N, int HW, int C,
                    void* stream) {
  return counted(launch_avgpool_fwd(H16(a_h), H16(a_l), e, N, HW, C, S(stream)));
}
int b2n_avgpool_bwd(const float* ge, const float* gate, float* g, int N, int HW, int C,
                    void* stream) {
  return counted(launch_avgpool_bwd(ge, gate, g, N, HW, C, S(stream)));
}

int b2n_linear_fwd(const float* x, long long ldx, const float* w, long long ldw, const float* b,
                   float* y, long long ldy, int rows, int in_f, int out_f, int relu, int accumulate,
                   void* stream) {
  return counted(launch_linear_fwd(x, ldx, w, ldw, b, y, ldy, rows, in_f, out_f, relu, accumulate,
                           S(stream)));
}
int b2n_linear_bwd_data(const float* dy, long long lddy, const float* w, long long ldw, float* dx,
                        long long lddx, const float* relu_mask, int rows, int in_f, int out_f,
                        int accumulate, void* stream) {
  return counted(launch_linear_bwd_data(dy, lddy, w, ldw, dx, lddx, relu_mask, rows, in_f, out_f,
                                accumulate, S(stream)));
}
int b2n_linear_bwd_weight(const float* dy, long long lddy, const float* x, long long ldx, float* dw,
                          long long lddw, float* db, int rows, int in_f, int out_f, int accumulate,
                          void* stream) {
  int rc = counted(launch_linear_bwd_weight(dy, lddy, x, ldx, dw, lddw, rows, in_f, out_f,
                                            accumulate, S(stream)));
  if (rc == 0 && db != nullptr)
    rc = counted(launch_colsum(dy, lddy, db, rows, out_f, accumulate, S(stream)));
  return rc;
}

int b2n_cols_replicate(float* y, long long ld, int rows, int width, int copies, void* stream) {
  return counted(launch_cols_replicate(y, ld, rows, width, copies, S(stream)));
}
int b2n_cols_sum(const float* dy, long long ld, float* out, int rows, int width, int copies,
                 void* stream) {
  return counted(launch_cols_sum(dy, ld, out, rows, width, copies, S(stream)));
}

int b2n_fused_loss(int mode, const float* logits_x, const long long* targets_i,
                   const float* targets_f, const float* logits_u_w, const float* logits_u_s,
                   int rows_x, int rows_u, int C, float lambda_u, float* losses, float* dlogits_x,
                   float* dlogits_u, long long* argmax_x, long long* pseudo_labels, void* stream) {
  return counted(launch_fused_loss(mode, logits_x, targets_i, targets_f, logits_u_w, logits_u_s, rows_x,
                           rows_u, C, lambda_u, losses, dlogits_x, dlogits_u, argmax_x,
                           pseudo_labels, S(stream)));
}

int b2n_softmax_last(const float* logits, float* out, int rows, int C, void* stream) {
  return counted(launch_softmax_last(logits, out, rows, C, S(stream)));
}

int b2n_lerp_multi(float* const* dst, float* const* src, const long long* numel, int n,
                   float alpha, int write_back, void* stream) {
  if (n < 0 || (n > 0 && (!dst || !src || !numel))) return set_error("b2n_lerp_multi: bad args");
  return counted(launch_lerp_multi(dst, src, numel, n, alpha, write_back, S(stream)),
                 (n + 95) / 96);
}

int b2n_adam_multi(float* const* p, const float* const* g, float* const* exp_avg,
                   float* const* exp_avg_sq, const long long* numel, int n, double lr, double beta1,
                   double beta2, double eps, double weight_decay, long long step,
                   long long* step_dev, double grad_scale, void* stream) {
  if (n < 0 || (n > 0 && (!p || !g || !exp_avg || !exp_avg_sq || !numel)))
    return set_error("b2n_adam_multi: bad args");
  return counted(launch_adam_multi(p, g, exp_avg, exp_avg_sq, numel, n, lr, beta1, beta2, eps,
                                   weight_decay, step, step_dev, grad_scale, S(stream)),
                 (n + 63) / 64 + (step_dev != nullptr ? 1 : 0));
}

int b2n_sgd_multi(float* const* p, const float* const* g, float* const* momentum_buf,
                  const long long* numel, int n, double lr, double momentum, double weight_decay,
                  int nesterov, int first_step, double grad_scale, void* stream) {
  if (n < 0 || (n > 0 && (!p || !g || !numel))) return set_error("b2n_sgd_multi: bad args");
  return counted(launch_sgd_multi(p, g, momentum_buf, numel, n, lr, momentum, weight_decay, nesterov,
                                  first_step, grad_scale, S(stream)),
                 (n + 63) / 64);
}

// ---- GPU augmentation ----
int b2n_aug_flip_crop(const unsigned char* src, unsigned char* dst, const int* top, const int* left,
                      const int* flip, int N, int Hs, int Ws, int H, int W, void* stream) {
  if (!src || !dst || !top || !left || !flip) return set_error("b2n_aug_flip_crop: null argument");
  return counted(launch_aug_flip_crop(src, dst, top, left, flip, N, Hs, Ws, H, W, S(stream)));
}
int b2n_aug_brightness_contrast(const unsigned char* src, unsigned char* dst, const float* alpha,
                                const float* offset, const int* apply, int N, int H, int W,
                                void* stream) {
  if (!src || !dst || !alpha || !offset) return set_error("b2n_aug_brightness_contrast: null argument");
  return counted(launch_aug_brightness_contrast(src, dst, alpha, offset, apply, N, H, W, S(stream)));
}
int b2n_aug_image_mean(const unsigned char* src, float* mean, int N, int H, int W, void* stream) {
  if (!src || !mean) return set_error("b2n_aug_image_mean: null argument");
  return counted(launch_aug_image_mean(src, mean, N, H, W, S(stream)));
}
int b2n_aug_hsv_shift(const unsigned char* src, unsigned char* dst, const int* dh, const int* ds,
                      const int* dv, const int* apply, int N, int H, int W, void* stream) {
  if (!src || !dst || !dh || !ds || !dv) return set_error("b2n_aug_hsv_shift: null argument");
  return counted(launch_aug_hsv_shift(src, dst, dh, ds, dv, apply, N, H, W, S(stream)));
}
int b2n_aug_add_noise(const unsigned char* src, unsigned char* dst, const float* noise, const int* apply,
                      int N, int H, int W, void* stream) {
  if (!src || !dst || !noise) return set_error("b2n_aug_add_noise: null argument");
  return counted(launch_aug_add_noise(src, dst, noise, apply, N, H, W, S(stream)));
}
int b2n_aug_box_blur(const unsigned char* src, unsigned char* dst, const int* ksize, const int* apply,
                     int N, int H, int W, void* stream) {
  if (!src || !dst || !ksize) return set_error("b2n_aug_box_blur: null argument");
  if (src == dst) return set_error("b2n_aug_box_blur: in-place operation is not supported");
  return counted(launch_aug_box_blur(src, dst, ksize, apply, N, H, W, S(stream)));
}
int b2n_aug_hed_jitter(const unsigned char* src, unsigned char* dst, const float* delta, const int* apply,
                       int N, int H, int W, void* stream) {
  if (!src || !dst || !delta) return set_error("b2n_aug_hed_jitter: null argument");
  return counted(launch_aug_hed_jitter(src, dst, delta, apply, N, H, W, S(stream)));
}
int b2n_aug_warp_affine(const unsigned char* src, unsigned char* dst, const float* minv, const int* apply,
                        int N, int Hs, int Ws, int H, int W, int clamp_border, void* stream) {
  if (!src || !dst || !minv) return set_error("b2n_aug_warp_affine: null argument");
  if (src == dst) return set_error("b2n_aug_warp_affine: in-place operation is not supported");
  return counted(launch_aug_warp_affine(src, dst, minv, apply, N, Hs, Ws, H, W, clamp_border, S(stream)));
}

}  // extern "C"
