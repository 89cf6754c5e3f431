/* b2n.h -- C ABI of libb2n.so: the B200-native (sm_100a) kernels behind the ResNet18
 * RSP-pretext / consistency-training hot path of srinidhiPY/SSL_CR_Histo.
 *
 * Conventions
 *   - every pointer is a CUDA device pointer owned by the caller (PyTorch's caching allocator
 *     in the Python binding); the library allocates nothing and keeps no hidden stream:
 *     `stream` is the caller's cudaStream_t (0 = legacy default stream); calls are async.
 *   - activations are NHWC fp32; tensors that feed a tensor-core conv hold TF32-representable
 *     values (the producing kernel rounds to nearest).  Parameters keep PyTorch layouts.
 *   - return 0 on success, non-zero on error; b2n_last_error() gives the message (thread-local).
 *     No exceptions cross the ABI.  The Python binding raises RuntimeError.
 *   - re-entrant; the only process-wide state is cached function attributes / SM count.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * repository; "tv:" = site-packages/torchvision/models/resnet.py, the third-party module that
 * models/net.py:32 instantiates).
 */
#ifndef B2N_H_
#define B2N_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2N_ABI_VERSION 2

/* IEEE binary16 storage (the FP16 operand planes of the forward convolutions). */
typedef uint16_t b2n_half;

int b2n_version(void);
const char* b2n_last_error(void);
/* kernels enqueued through this library since it was loaded (monotonic, process-wide). */
unsigned long long b2n_launch_count(void);
/* 1 when the current device is compute capability 10.x, 0 otherwise, <0 when no device. */
int b2n_device_ok(void);

/* ---- tensor-core convolutions ---------------------------------------------------------- */

/* Implicit-GEMM convolution on tcgen05 tensor cores, FP32 accumulate in TMEM.
 *   y[n,p,q,k] = sum_{r,s,c} x[n, p*stride - pad_h_lo + r, q*stride - pad_w_lo + s, c]
 *                             * w[k][(r*S+s)*Cin + c]
 * Two operand modes:
 *   - error-compensated forward mode (x_h, x_l, w_h, w_l given; x = w_packed = NULL): both
 *     operands are (hi, lo) pairs of FP16 tensors, hi = fp16(v), lo = fp16(v - hi); the kernel
 *     accumulates hi*hi + lo*hi + hi*lo (kind::f16) -- FP32-grade results.  Cin = 32 or a
 *     multiple of 64; Cin = 16 for stride-1 same-width convs with Cout = 64 (the stem).
 *   - plain mode (x, w_packed given; halves NULL): one TF32 pass over fp32 containers (data
 *     gradients, with a b2n_pack_weight_dgrad pack).  Cin a multiple of 32.
 * Stride-1 same-width convs with Cout = 64 whose packed weights fit in shared memory (layer1's
 * 3x3 convs, the 4x4 stem) run a tap-sharing variant: one activation box per filter row, the
 * horizontal taps read through row-shifted UMMA descriptors (same results, ~3x less L2 traffic).
 * epilogue, in this order: v = acc; v = v*scale[k] + shift[k] (scale and shift come as a pair);
 *           v += resid[..] (fp32; only where mask[..] > 0 if mask); v += resid_h + resid_l (FP16
 *           pair); relu;
 *           (data-gradient launches, both optional) v = 0 where gate[..] <= 0 (the ReLU gate of
 *           the activation this gradient belongs to; addressed like the output; not together with
 *           mask); and -- bnb_y given -- v is treated as the gradient g w.r.t. relu(bn(y)) of the
 *           BatchNorm whose raw input is bnb_y: with bnb_scale/shift the gate
 *           fmaf(y, scale, shift) > 0 is applied to g, and stats[0][k] += sum g,
 *           stats[1][k] += sum g * (y - bnb_mean[k]) * bnb_invstd[k] (b2n_bn_bwd_reduce's sums,
 *           taken while the tile is still on chip; dense fp32 result only);
 *           finally y (fp32, TF32-rounded if round_tf32) and / or the (hi, lo) FP16 pair
 *           y_h / y_l are stored.
 * Output placement: o_step == 0 -> dense [N,P,Q,Cout]; otherwise output pixel (i, j) of image n
 * goes to (o_h0 + i*o_step, o_w0 + j*o_step) of an [N,o_H,o_W,Cout] tensor (resid / mask are
 * addressed the same way; pixels outside are dropped) -- the parity classes of a stride-2 data
 * gradient (b2n_pack_weight_dgrad_s2).
 * stats (optional, [2][Cout] doubles, caller-zeroed; not together with scale/shift): += per-channel
 * sum / sum of squares of the raw accumulator -- the BatchNorm batch statistics (without bnb_y).
 * Cout a multiple of 64.
 * Replaces: tv:92,96,100 (conv3x3 / downsample conv in BasicBlock.forward), tv:268 (stem, via
 * the space-to-depth view), and the conv dgrad reached from loss.backward()
 * (pretrain_BreastPathQ.py:60, eval_BreastPathQ_SSL_CR.py:99). */
int b2n_conv_fwd(const float* x, const b2n_half* x_h, const b2n_half* x_l, const float* w_packed,
                 const b2n_half* w_h, const b2n_half* w_l, float* y, b2n_half* y_h, b2n_half* y_l,
                 int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad_h_lo,
                 int pad_h_hi, int pad_w_lo, int pad_w_hi, const float* scale, const float* shift,
                 const float* resid, const b2n_half* resid_h, const b2n_half* resid_l,
                 const float* mask, int relu, int round_tf32, double* stats,
                 const int* x_l_nonzero /* optional device flag: 0 => x_l is all zero, skip it */,
                 int o_step, int o_h0, int o_w0, int o_H, int o_W, const float* gate,
                 const float* bnb_y, const float* bnb_mean, const float* bnb_invstd,
                 const float* bnb_scale, const float* bnb_shift, void* stream);

/* Weight gradient, split-K over pixels, accumulated atomically:
 *   dw_packed[k][(r*S+s)*Cin + c] += sum_{n,p,q} dy[n,p,q,k] * x[n, p*stride-pad+r, q*stride-pad+s, c]
 * dw_packed must be zeroed by the caller.  Cin a multiple of 32, Cout a multiple of 64.  Stride-1
 * same-width convs with Cout = 64 and R*Cin/32 <= 6 (layer1, the stem) run the tap-sharing kernel
 * in which one CTA owns all of dW for a slab of pixels.  Replaces cudnnConvolutionBackwardFilter reached from
 * loss.backward() (pretrain_BreastPathQ.py:60).
 * deterministic != 0: no atomics -- every split-K CTA group stores its own partial plane, so
 * dw_packed must hold b2n_conv_wgrad_planes(...) planes of [Cout][R*S*Cin] floats (no zero-fill
 * needed) and b2n_unpack_wgrad sums them in a fixed order: bit-repeatable gradients.
 * x_channels (0 = Cin): channels x actually stores per pixel; with Cin = 32 it may be smaller (a
 * multiple of 4): the remaining reduction channels are zero-filled by the TMA unit instead of being
 * stored -- the stem's space-to-depth input has 12 real channels. */
int b2n_conv_wgrad(const float* x, const float* dy, float* dw_packed, int N, int H, int W, int Cin,
                   int Cout, int R, int S, int stride, int pad_h_lo, int pad_h_hi, int pad_w_lo,
                   int pad_w_hi, int deterministic, int x_channels, void* stream);
/* Planes a deterministic b2n_conv_wgrad of this shape writes on the current device (>= 1; < 0 on
 * error).  Host-side only, no launch. */
int b2n_conv_wgrad_planes(int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                          int pad_h_lo, int pad_h_hi, int pad_w_lo, int pad_w_hi);

/* (K,C,R,S) parameter -> forward pack [K][(r*S+s)*C + c] as a (hi, lo) FP16 pair, data-gradient
 * pack [C][((R-1-r)*S+(S-1-s))*K + k] (TF32-rounded fp32); packed weight gradient -> (K,C,R,S). */
int b2n_pack_weight_fwd(const float* w, b2n_half* w_h, b2n_half* w_l, int K, int C, int R, int S,
                        void* stream);
int b2n_pack_weight_dgrad(const float* w, float* w_packed, int K, int C, int R, int S, void* stream);
/* n packs in one launch (64 per launch): HOST arrays of device pointers and shapes, as
 * b2n_lerp_multi.  kind[i] = 0: b2n_pack_weight_fwd of w[i] into (dst0[i], dst1[i]) = (hi, lo);
 * 1: b2n_pack_weight_dgrad into dst0[i]; 2: b2n_pack_weight_dgrad_s2m into dst0[i] (3x3 only).
 * A training step repacks every conv weight of the trunk after its optimizer step
 * (pretrain_BreastPathQ.py:61): ~40 packs of a few microseconds each. */
int b2n_pack_weights_multi(const float* const* w, void* const* dst0, void* const* dst1, const int* kind,
                           const int* K, const int* C, const int* R, const int* S, int n, void* stream);
/* Stride-2 3x3/pad-1 data gradient by output parity: four stride-1 tap subsets (1, 2, 2, 4 taps)
 * over dY, packs stored back to back [C][ntaps*K] in class order (0,0), (0,1), (1,0), (1,1);
 * 9*C*K floats. */
int b2n_pack_weight_dgrad_s2(const float* w, float* w_packed, int K, int C, void* stream);
/* The same nine (class, tap) blocks as one K-major matrix [C][9*K], and the merged launch that uses
 * it: all four parity classes of the stride-2 3x3/pad-1 data gradient from ONE pass over
 * dy [N,P,Q,K] (four TMEM accumulators per 128-pixel tile):
 *   dx[n, 2i+ph, 2j+pw, c] = sum_{a <= ph, b <= pw, k} dy[n, i+a, j+b, k] * w[k][c][rr(ph,a)][ss(pw,b)]
 * dx is [N,H,W,C] with P = ceil(H/2), Q = ceil(W/2); resid (optional, may alias dx) is added on the
 * even-even pixels (the 1x1 shortcut conv's gradient); gate (optional) zeroes dx where gate <= 0.
 * K a multiple of 32, C of 64.  Replaces cudnnConvolutionBackwardData of the three stride-2 convs
 * (tv:92 in layer{2,3,4}.0) reached from loss.backward(). */
int b2n_pack_weight_dgrad_s2m(const float* w, float* w_packed, int K, int C, void* stream);
int b2n_conv_dgrad_s2(const float* dy, const float* w_packed, float* dx, int N, int P, int Q, int K,
                      int C, int H, int W, const float* resid, const float* gate, void* stream);
/* The same with the block's 1x1 stride-2 shortcut conv folded in (torchvision BasicBlock.downsample,
 * tv:100-101): its data gradient only reaches the even-even pixels, so dy_sc [N,P,Q,K] times the
 * shortcut's b2n_pack_weight_dgrad pack w_sc_packed [C][K] is accumulated into that parity class
 * inside the kernel -- no separate 1x1 launch and no resid round trip through dx. */
int b2n_conv_dgrad_s2_sc(const float* dy, const float* w_packed, const float* dy_sc, const float* w_sc_packed,
                         float* dx, int N, int P, int Q, int K, int C, int H, int W, const float* gate,
                         void* stream);
/* accumulate != 0: dw += (several passes over shared weights feed one gradient slot, e.g. the
 * three trunk passes of TripletNet.forward or a flat all-reduce arena). */
int b2n_unpack_wgrad(const float* dw_packed, float* dw, int K, int C, int R, int S, int accumulate,
                     int planes /* partial planes to sum, 1 = plain */, void* stream);

/* ---- stem: 7x7/s2 conv as a 4x4/s1 conv over a 2x2 space-to-depth view (tv:197,268) ----- */
/* x NCHW fp32 (N,3,H,W), H and W even -> NHWC space-to-depth views (12 real channels): the
 * (hi, lo) FP16 pair (N,H/2,W/2,16) for the forward conv and (xs32, may be NULL) the TF32 fp32
 * copy (N,H/2,W/2,12) for the wgrad (b2n_conv_wgrad with Cin = 32, x_channels = 12). */
int b2n_stem_pack_input(const float* x_nchw, b2n_half* xs_h, b2n_half* xs_l, float* xs32,
                        int* xs_l_nonzero /* optional, caller-zeroed: set to 1 if any lo != 0 */,
                        int N, int H, int W, void* stream);
/* Same for uint8 pixels (the patches as the dataset holds them, dataset.py:65-67, before the
 * loop's .float()): 4x fewer host->device bytes.  uint8 values are exact in FP16, so there is no
 * lo plane: pass the conv a zeroed x_l_nonzero flag and any valid pointer as x_l. */
int b2n_stem_pack_input_u8(const unsigned char* x_nchw, b2n_half* xs_h, float* xs32, int N, int H,
                           int W, void* stream);
/* w (K,3,7,7) -> (hi, lo) FP16 pair [K][16 taps * 16];  packed gradient [K][16 taps * 32] ->
 * (K,3,7,7). */
int b2n_stem_pack_weight(const float* w, b2n_half* ws_h, b2n_half* ws_l, int K, void* stream);
int b2n_stem_unpack_wgrad(const float* dws, float* dw, int K, int accumulate, int planes,
                          void* stream);

/* ---- BatchNorm / ReLU / residual (tv:93-103,269-270; nn.BatchNorm2d train + eval) -------- */
/* Batch statistics -> per-channel affine (scale = gamma*invstd, shift = beta - mean*scale),
 * saves mean / invstd for backward, and applies `n_updates` running-stat updates in closed form
 * (momentum, unbiased variance).  running_* may be null (no update). */
int b2n_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float* scale, float* shift, float* mean, float* invstd,
                    float* inv_gamma /* optional: 1/gamma (0 where gamma == 0) */, int C, double count,
                    float momentum, float eps, int n_updates, void* stream);
/* eval mode: scale = gamma / sqrt(running_var + eps), shift = beta - running_mean*scale */
int b2n_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean,
                     const float* running_var, float* scale, float* shift, int C, float eps,
                     void* stream);
/* the same for n layers in one launch (32 per launch; HOST arrays of device pointers, channel
 * counts and epsilons): all twenty BatchNorm layers of an eval-mode trunk pass */
int b2n_bn_fold_eval_multi(const float* const* gamma, const float* const* beta,
                           const float* const* running_mean, const float* const* running_var,
                           float* const* scale, float* const* shift, const int* C, const float* eps,
                           int n, void* stream);
/* v = [relu](scale*y + shift + residual); residual = res32, or res_scale*res32 + res_shift when
 * res_scale is given (downsample-branch BN), or the FP16 pair res_h + res_l (identity shortcut).
 * Outputs, each optional: out32 (fp32, TF32-rounded if round_tf32: the backward pass' operand and
 * ReLU mask) and the (hi, lo) FP16 pair out_h / out_l (the next forward conv's operand).
 * [rows][C] row-major, C a multiple of 8. */
int b2n_bn_apply(const float* y, const float* scale, const float* shift, const float* res32,
                 const float* res_scale, const float* res_shift, const b2n_half* res_h,
                 const b2n_half* res_l, float* out32, b2n_half* out_h, b2n_half* out_l,
                 long long rows, int C, int relu, int round_tf32, void* stream);
/* BatchNorm backward, two passes.  g' = g * [ReLU gate]; the gate is mask > 0 (mask = post-ReLU
 * block output), or -- when the ReLU input is this BN's own output (bn1 of a BasicBlock) --
 * fmaf(y, gate_scale, gate_shift) > 0 recomputed from y with the forward affine (saves reading the
 * mask in both passes), or absent (all three null).
 *   reduce: sums[0][c] += sum g', sums[1][c] += sum g' * xhat        (sums caller-zeroed)
 *   apply : dy = gamma*invstd*(g' - sums0/rows - xhat*sums1/rows); dgamma = sums1, dbeta = sums0 */
int b2n_bn_bwd_reduce(const float* g, const float* mask, const float* y, const float* mean,
                      const float* invstd, const float* gate_scale, const float* gate_shift,
                      double* sums, long long rows, int C, void* stream);
int b2n_bn_bwd_apply(const float* g, const float* mask, const float* y, const float* mean,
                     const float* invstd, const float* gamma, const float* gate_scale,
                     const float* gate_shift, const double* sums, float* dy, float* dgamma,
                     float* dbeta, long long rows, int C, int round_tf32,
                     int accumulate /* dgamma / dbeta += instead of = */, void* stream);
/* up[n,2p,2q,:] = dy[n,p,q,:], zero elsewhere: stride-2 data gradient as a stride-1 conv. */
int b2n_upsample_zero(const float* dy, float* up, int N, int P, int Q, int H, int W, int C,
                      void* stream);

/* ---- pooling (tv:271 MaxPool2d(3,2,1) fused with bn1+relu; tv:278-279 avgpool+flatten) --- */
int b2n_bn_relu_maxpool(const float* y, const float* scale, const float* shift,
                        float* a32 /* may be null */, b2n_half* a_h, b2n_half* a_l,
                        unsigned char* argmax_idx /* may be null */, int N, int H, int W, int C,
                        void* stream);
int b2n_maxpool_relu_bwd(const float* ga, const unsigned char* argmax_idx, const float* y,
                         const float* scale, const float* shift, float* gz, int N, int H, int W,
                         int C, void* stream);
/* The stem's tail backward in two sweeps instead of five: the gradient of maxpool+ReLU is rebuilt
 * on the fly from the pooled gradient `ga` (N,P,Q,C) and the recorded argmax, and fed straight
 * into the two BatchNorm-backward passes (same formulas as b2n_bn_bwd_reduce / _apply; y is the
 * raw stem conv output (N,H,W,C), scale/shift the forward BN affine that defines the ReLU gate).
 * Replaces loss.backward() through tv:269-271 (bn1, relu, maxpool). */
int b2n_pool_bn_bwd_reduce(const float* ga, const unsigned char* argmax_idx, const float* y,
                           const float* scale, const float* shift, const float* mean,
                           const float* invstd, double* sums /* [2][C], caller-zeroed */, int N,
                           int H, int W, int C, void* stream);
int b2n_pool_bn_bwd_apply(const float* ga, const unsigned char* argmax_idx, const float* y,
                          const float* scale, const float* shift, const float* mean,
                          const float* invstd, const float* gamma, const double* sums, float* dy,
                          float* dgamma, float* dbeta, int N, int H, int W, int C, int round_tf32,
                          int accumulate /* dgamma / dbeta += instead of = */, void* stream);
int b2n_avgpool_fwd(const b2n_half* a_h, const b2n_half* a_l, float* e, int N, int HW, int C,
                    void* stream);
/* g[n,i,c] = ge[n,c] / HW, zeroed where gate[n,i,c] <= 0 (gate optional: the pooled activation,
 * whose ReLU gate is applied here so that no consumer of g reads a mask). */
int b2n_avgpool_bwd(const float* ge, const float* gate, float* g, int N, int HW, int C,
                    void* stream);

/* ---- fully-connected heads, exact FP32 (models/net.py:12-15,36-37,60-62,110) ------------- */
int b2n_linear_fwd(const float* x, long long ldx, const float* w, long long ldw, const float* b,
                   float* y, long long ldy, int rows, int in_f, int out_f, int relu, int accumulate,
                   void* stream);
int b2n_linear_bwd_data(const float* dy, long long lddy, const float* w, long long ldw, float* dx,
                        long long lddx, const float* relu_mask, int rows, int in_f, int out_f,
                        int accumulate, void* stream);
int b2n_linear_bwd_weight(const float* dy, long long lddy, const float* x, long long ldx, float* dw,
                          long long lddw, float* db /* may be null */, int rows, int in_f,
                          int out_f, int accumulate, void* stream);

/* y[r][k*width + c] = y[r][c] for k = 1..copies-1, and out[r][c] = sum_k dy[r][k*width + c]: the
 * pair features cat(f, f, f) of TripletNet_Finetune.forward (models/net.py:92-103) and their
 * gradient, when the three trunk passes saw the same input (one evaluation of the pair MLP). */
int b2n_cols_replicate(float* y, long long ld, int rows, int width, int copies, void* stream);
int b2n_cols_sum(const float* dy, long long ld, float* out, int rows, int width, int copies,
                 void* stream);

/* ---- fused losses ------------------------------------------------------------------------
 * mode 0: mean softmax-CE of logits_x vs targets_i + argmax (pretrain_BreastPathQ.py:56,66)
 * mode 1: sup = MSE(logits_x, targets_f), cons = MSE(logits_u_w, logits_u_s)
 *         (eval_BreastPathQ_SSL_CR.py:92-95)
 * mode 2: sup = CE(logits_x, targets_i), cons = CE(logits_u_s, argmax logits_u_w)
 *         (eval_Kather_SSL_CR.py:87-93)
 * losses[3] = {sup, cons, sup + lambda_u*cons}; dlogits_* = d(total)/d(logits) (may be null). */
int b2n_fused_loss(int mode, const float* logits_x, const long long* targets_i,
                   const float* targets_f, const float* logits_u_w, const float* logits_u_s,
                   int rows_x, int rows_u, int C, float lambda_u, float* losses, float* dlogits_x,
                   float* dlogits_u, long long* argmax_x, long long* pseudo_labels, void* stream);

/* out[i] = softmax(logits[i, :])[C-1]: the tumour-probability column of the WSI heat-map
 * inference loop (test_Camelyon16.py:57-58). */
int b2n_softmax_last(const float* logits, float* out, int rows, int C, void* stream);

/* ---- multi-tensor weight lerp ("EMA") -----------------------------------------------------
 * dst[i] <- alpha*src[i] + (1-alpha)*dst[i]; write_back also stores the result into src[i].
 * dst/src/numel are HOST arrays of n device pointers / element counts.
 * alpha=1: teacher<-student hand-off (eval_BreastPathQ_SSL_CR.py:515-516);
 * alpha=1-la_alpha, write_back=1: Lookahead pull (models/optimiser/RAdam/lookahead.py:96-97). */
int b2n_lerp_multi(float* const* dst, float* const* src, const long long* numel, int n,
                   float alpha, int write_back, void* stream);

/* ---- multi-tensor optimizer steps (torch.optim formulas term by term) -- */
/* Host arrays of n device pointers / element counts (as b2n_lerp_multi).  grad_scale multiplies
 * every gradient first (1/world after a summing all-reduce, else 1).
 * Adam with L2 weight decay, bias correction for `step` >= 1:
 *   replaces optimizer.step() of torch.optim.Adam (eval_BreastPathQ_SSL_CR.py:481,100;
 *   eval_Kather_SSL.py:419,72). */
int b2n_adam_multi(float* const* p, const float* const* g, float* const* exp_avg,
                   float* const* exp_avg_sq, const long long* numel, int n, double lr, double beta1,
                   double beta2, double eps, double weight_decay, long long step,
                   long long* step_dev /* optional device counter: when given it is incremented and
                   used instead of `step`, so the launch can be replayed from a CUDA graph */,
                   double grad_scale, void* stream);
/* SGD with momentum (dampening 0), optional Nesterov, L2 weight decay; first_step != 0
 * initialises the momentum buffer with the gradient:
 *   replaces torch.optim.SGD(nesterov=True) (pretrain_BreastPathQ.py:245,61;
 *   eval_Camelyon_SSL_CR.py:514). */
int b2n_sgd_multi(float* const* p, const float* const* g, float* const* momentum_buf,
                  const long long* numel, int n, double lr, double momentum, double weight_decay,
                  int nesterov, int first_step, double grad_scale, void* stream);

/* ---- weak / strong augmentation on the GPU --------------------------------
 * Batches are uint8 (N,3,H,W), planes contiguous -- the layout the trunk's uint8 input path takes.
 * Per-image parameters are device arrays of N entries; `apply` (may be NULL = all) selects the
 * images an operation touches, the others are copied through.  src == dst is allowed for the
 * point-wise operations (not for blur / warp).  Replaces, one launch per batch instead of Python
 * per image: TransformFix (dataset.py:663-677) and the RandAugment pool
 * (models/randaugment.py:51-126), whose arithmetic lives in albumentations 0.1.8 / imgaug 0.4 /
 * scikit-image 0.15 / OpenCV (requirements.txt:10,128,369,241). */
/* RandomHorizontalFlip + RandomCrop: dst[n,c,y,x] = flipped(src[n])[c, top+y, left+x]. */
int b2n_aug_flip_crop(const unsigned char* src, unsigned char* dst, const int* top, const int* left,
                      const int* flip, int N, int Hs, int Ws, int H, int W, void* stream);
/* albumentations brightness_contrast_adjust: uint8(clip(float32(x) * alpha[n] + offset[n], 0, 255)). */
int b2n_aug_brightness_contrast(const unsigned char* src, unsigned char* dst, const float* alpha,
                                const float* offset, const int* apply, int N, int H, int W,
                                void* stream);
/* mean[n] = mean over the image's 3*H*W bytes (the `beta * np.mean(img)` brightness reference). */
int b2n_aug_image_mean(const unsigned char* src, float* mean, int N, int H, int W, void* stream);
/* albumentations shift_hsv (HueSaturationValue): OpenCV 8-bit RGB->HSV, integer shifts, HSV->RGB. */
int b2n_aug_hsv_shift(const unsigned char* src, unsigned char* dst, const int* dh, const int* ds,
                      const int* dv, const int* apply, int N, int H, int W, void* stream);
/* imgaug AdditiveGaussianNoise(per_channel=False): uint8(clip(round(x + noise[n,0,y,x]))). */
int b2n_aug_add_noise(const unsigned char* src, unsigned char* dst, const float* noise, const int* apply,
                      int N, int H, int W, void* stream);
/* albumentations Blur = cv2.blur: ksize[n] x ksize[n] box filter (odd, <= 7), BORDER_REFLECT_101. */
int b2n_aug_box_blur(const unsigned char* src, unsigned char* dst, const int* ksize, const int* apply,
                     int N, int H, int W, void* stream);
/* colour_augmentation (models/randaugment.py:17-48): rgb2hed, stain offsets delta[n][3], hed2rgb. */
int b2n_aug_hed_jitter(const unsigned char* src, unsigned char* dst, const float* delta, const int* apply,
                       int N, int H, int W, void* stream);
/* cv2.warpAffine / cv2.resize with INTER_CUBIC: minv[n][6] maps output (x, y) to source (sx, sy);
 * BORDER_REFLECT_101, or edge replication when clamp_border (resize).  (N,3,Hs,Ws) -> (N,3,H,W). */
int b2n_aug_warp_affine(const unsigned char* src, unsigned char* dst, const float* minv, const int* apply,
                        int N, int Hs, int Ws, int H, int W, int clamp_border, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2N_H_ */
