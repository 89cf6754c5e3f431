// BatchNorm + ReLU + residual element-wise kernels (NHWC fp32, HBM-bound, 128-bit accesses).
//
// They replace cudnnBatchNormalizationForwardTraining/Backward and the ATen ReLU / add kernels
// behind torchvision BasicBlock.forward (site-packages/torchvision/models/resnet.py:93-103).
// The per-channel batch statistics themselves (sum, sum of squares) come out of the conv
// kernel's epilogue (conv_igemm.cuh); here they are finalised, applied, and differentiated.
#include "launch.h"
#include "ptx.cuh"

namespace b2n {

// ------------------------------------------------------------------ finalize
// stats = [2][C] doubles (sum, sumsq) over `count` values per channel.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, int C, double count,
                                   float momentum, float eps, int n_updates) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = stats[c] / count;
  double var = stats[C + c] / count - mean * mean;
  if (var < 0) var = 0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - static_cast<float>(mean) * sc;
  mean_out[c] = static_cast<float>(mean);
  invstd_out[c] = invstd;
  if (running_mean != nullptr && n_updates > 0) {
    // n_updates identical updates r <- (1-m) r + m b collapse to one closed-form update;
    // n_updates = 3 reproduces TripletNet_Finetune.forward's three passes (models/net.py:88-90).
    const double keep = pow(1.0 - static_cast<double>(momentum), n_updates);
    const double unbiased = count > 1 ? var * count / (count - 1.0) : var;
    running_mean[c] = static_cast<float>(keep * running_mean[c] + (1.0 - keep) * mean);
    running_var[c] = static_cast<float>(keep * running_var[c] + (1.0 - keep) * unbiased);
  }
}

// eval mode: fold running statistics into a per-channel affine
__global__ void bn_fold_eval_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ rm, const float* __restrict__ rv,
                                    float* __restrict__ scale, float* __restrict__ shift, int C,
                                    float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}

// --------------------------------------------------------------------- apply
// out = act(scale*y + shift + residual), residual = res (identity) or res_scale*res + res_shift
// (the downsample branch's BN), optionally rounded to TF32 because `out` feeds the next conv.
// OUT_MODE: 0 = plain fp32, 1 = TF32-rounded, 2 = (hi, lo) TF32 pair into out / out_lo.
// res_lo: low part when the residual is itself a (hi, lo) pair (identity shortcut).
template <bool RELU, int OUT_MODE>
__global__ void bn_apply_kernel(const float4* __restrict__ y, const float* __restrict__ scale,
                                const float* __restrict__ shift, const float4* __restrict__ res,
                                const float4* __restrict__ res_lo,
                                const float* __restrict__ res_scale,
                                const float* __restrict__ res_shift, float4* __restrict__ out,
                                float4* __restrict__ out_lo, size_t n4, int C) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int c = static_cast<int>((i * 4) % C);
    const float4 v = y[i];
    const float4 sc = *reinterpret_cast<const float4*>(scale + c);
    const float4 sh = *reinterpret_cast<const float4*>(shift + c);
    float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z),
                           fmaf(v.w, sc.w, sh.w));
    if (res != nullptr) {
      float4 r = res[i];
      if (res_lo != nullptr) {
        const float4 rl = res_lo[i];
        r.x += rl.x; r.y += rl.y; r.z += rl.z; r.w += rl.w;
      }
      if (res_scale != nullptr) {
        const float4 rs = *reinterpret_cast<const float4*>(res_scale + c);
        const float4 rh = *reinterpret_cast<const float4*>(res_shift + c);
        r = make_float4(fmaf(r.x, rs.x, rh.x), fmaf(r.y, rs.y, rh.y), fmaf(r.z, rs.z, rh.z),
                        fmaf(r.w, rs.w, rh.w));
      }
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (RELU) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    if (OUT_MODE == 2) {
      const float4 h = make_float4(tf32_rn(o.x), tf32_rn(o.y), tf32_rn(o.z), tf32_rn(o.w));
      out[i] = h;
      out_lo[i] = make_float4(tf32_rn(o.x - h.x), tf32_rn(o.y - h.y), tf32_rn(o.z - h.z),
                              tf32_rn(o.w - h.w));
    } else {
      if (OUT_MODE == 1) {
        o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w);
      }
      out[i] = o;
    }
  }
}

int launch_bn_apply(const float* y, const float* scale, const float* shift, const float* res,
                    const float* res_lo, const float* res_scale, const float* res_shift,
                    float* out, float* out_lo, long long rows, int C, int relu, int round_tf32,
                    cudaStream_t stream) {
  if (C % 4 != 0) return set_error("bn_apply: C %% 4 != 0");
  const size_t n4 = static_cast<size_t>(rows) * C / 4;
  if (n4 == 0) return 0;
  const int threads = 256;
  size_t blocks = (n4 + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  auto y4 = reinterpret_cast<const float4*>(y);
  auto r4 = reinterpret_cast<const float4*>(res);
  auto rl4 = reinterpret_cast<const float4*>(res_lo);
  auto o4 = reinterpret_cast<float4*>(out);
  auto ol4 = reinterpret_cast<float4*>(out_lo);
  const int mode = out_lo != nullptr ? 2 : (round_tf32 ? 1 : 0);
#define B2N_LAUNCH(R, T)                                                                     \
  bn_apply_kernel<R, T><<<(unsigned)blocks, threads, 0, stream>>>(y4, scale, shift, r4, rl4, \
                                                                 res_scale, res_shift, o4, ol4, n4, C)
  if (relu) {
    if (mode == 2) B2N_LAUNCH(true, 2);
    else if (mode == 1) B2N_LAUNCH(true, 1);
    else B2N_LAUNCH(true, 0);
  } else {
    if (mode == 2) B2N_LAUNCH(false, 2);
    else if (mode == 1) B2N_LAUNCH(false, 1);
    else B2N_LAUNCH(false, 0);
  }
#undef B2N_LAUNCH
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("bn_apply: %s", cudaGetErrorString(e));
  return 0;
}

int launch_bn_finalize(const double* stats, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float* scale, float* shift,
                       float* mean, float* invstd, int C, double count, float momentum, float eps,
                       int n_updates, cudaStream_t stream) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(stats, gamma, beta, running_mean,
                                                         running_var, scale, shift, mean, invstd, C,
                                                         count, momentum, eps, n_updates);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("bn_finalize: %s", cudaGetErrorString(e));
  return 0;
}

int launch_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                        float* scale, float* shift, int C, float eps, cudaStream_t stream) {
  bn_fold_eval_kernel<<<(C + 127) / 128, 128, 0, stream>>>(gamma, beta, rm, rv, scale, shift, C,
                                                          eps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("bn_fold_eval: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------ backward
// g' = g * (mask > 0)   (mask = the block's post-ReLU output; null = no ReLU gate)
// sums[0][c] = sum g',  sums[1][c] = sum g' * xhat,  xhat = (y - mean) * invstd
//
// Thread layout: threadIdx.x -> float4 channel group (C/4 of them), threadIdx.y -> row lane.
__global__ void bn_bwd_reduce_kernel(const float4* __restrict__ g, const float4* __restrict__ mask,
                                     const float4* __restrict__ y, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, double* __restrict__ sums,
                                     long long rows, int C, int rows_per_block) {
  extern __shared__ float red[];  // [blockDim.y][C][2]
  const int cg = threadIdx.x;     // channel group
  const int C4 = C >> 2;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
  if (cg < C4) {
    const float4 mu = *reinterpret_cast<const float4*>(mean + 4 * cg);
    const float4 is = *reinterpret_cast<const float4*>(invstd + 4 * cg);
    for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      const size_t i = static_cast<size_t>(r) * C4 + cg;
      float4 gv = g[i];
      if (mask != nullptr) {
        const float4 m = mask[i];
        gv.x = m.x > 0.f ? gv.x : 0.f; gv.y = m.y > 0.f ? gv.y : 0.f;
        gv.z = m.z > 0.f ? gv.z : 0.f; gv.w = m.w > 0.f ? gv.w : 0.f;
      }
      const float4 yv = y[i];
      s1.x += gv.x; s1.y += gv.y; s1.z += gv.z; s1.w += gv.w;
      s2.x += gv.x * (yv.x - mu.x) * is.x; s2.y += gv.y * (yv.y - mu.y) * is.y;
      s2.z += gv.z * (yv.z - mu.z) * is.z; s2.w += gv.w * (yv.w - mu.w) * is.w;
    }
    float* dst = red + (static_cast<size_t>(threadIdx.y) * C + 4 * cg) * 2;
    dst[0] = s1.x; dst[1] = s2.x; dst[2] = s1.y; dst[3] = s2.y;
    dst[4] = s1.z; dst[5] = s2.z; dst[6] = s1.w; dst[7] = s2.w;
  }
  __syncthreads();
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int j = tid; j < 2 * C; j += blockDim.x * blockDim.y) {
    float acc = 0.f;
    for (int ry = 0; ry < blockDim.y; ++ry) acc += red[static_cast<size_t>(ry) * 2 * C + j];
    const int c = j >> 1, which = j & 1;
    atomicAdd(&sums[which * C + c], static_cast<double>(acc));
  }
}

// dy = gamma*invstd * (g' - sum_g/M - xhat * sum_gx/M); block 0 also emits dgamma / dbeta.
template <bool ROUND>
__global__ void bn_bwd_apply_kernel(const float4* __restrict__ g, const float4* __restrict__ mask,
                                    const float4* __restrict__ y, const float* __restrict__ mean,
                                    const float* __restrict__ invstd,
                                    const float* __restrict__ gamma,
                                    const double* __restrict__ sums, float4* __restrict__ dy,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, size_t n4,
                                    int C, double inv_count) {
  if (blockIdx.x == 0 && dgamma != nullptr) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dbeta[c] = static_cast<float>(sums[c]);
      dgamma[c] = static_cast<float>(sums[C + c]);
    }
  }
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int c = static_cast<int>((i * 4) % C);
    float4 gv = g[i];
    if (mask != nullptr) {
      const float4 m = mask[i];
      gv.x = m.x > 0.f ? gv.x : 0.f; gv.y = m.y > 0.f ? gv.y : 0.f;
      gv.z = m.z > 0.f ? gv.z : 0.f; gv.w = m.w > 0.f ? gv.w : 0.f;
    }
    const float4 yv = y[i];
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 is = *reinterpret_cast<const float4*>(invstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
    float o[4];
    const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
    const float mm[4] = {mu.x, mu.y, mu.z, mu.w};
    const float ii[4] = {is.x, is.y, is.z, is.w};
    const float aa[4] = {ga.x, ga.y, ga.z, ga.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float m1 = static_cast<float>(sums[c + k] * inv_count);
      const float m2 = static_cast<float>(sums[C + c + k] * inv_count);
      const float xh = (yy[k] - mm[k]) * ii[k];
      float v = aa[k] * ii[k] * (gg[k] - m1 - xh * m2);
      o[k] = ROUND ? tf32_rn(v) : v;
    }
    dy[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

int launch_bn_bwd_reduce(const float* g, const float* mask, const float* y, const float* mean,
                         const float* invstd, double* sums, long long rows, int C,
                         cudaStream_t stream) {
  if (C % 4 != 0 || C > 1024) return set_error("bn_bwd_reduce: unsupported C=%d", C);
  const int C4 = C / 4;
  int ty = 256 / C4;
  if (ty < 1) ty = 1;
  if (ty > 16) ty = 16;
  dim3 block(C4, ty);
  const int target_blocks = device_sm_count() * 4;
  long long rpb = (rows + target_blocks - 1) / target_blocks;
  if (rpb < 4LL * ty) rpb = 4LL * ty;
  const int blocks = static_cast<int>((rows + rpb - 1) / rpb);
  const size_t smem = static_cast<size_t>(ty) * C * 2 * sizeof(float);
  bn_bwd_reduce_kernel<<<blocks, block, smem, stream>>>(
      reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(mask),
      reinterpret_cast<const float4*>(y), mean, invstd, sums, rows, C, static_cast<int>(rpb));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("bn_bwd_reduce: %s", cudaGetErrorString(e));
  return 0;
}

int launch_bn_bwd_apply(const float* g, const float* mask, const float* y, const float* mean,
                        const float* invstd, const float* gamma, const double* sums, float* dy,
                        float* dgamma, float* dbeta, long long rows, int C, int round_tf32,
                        cudaStream_t stream) {
  if (C % 4 != 0) return set_error("bn_bwd_apply: C %% 4 != 0");
  const size_t n4 = static_cast<size_t>(rows) * C / 4;
  const int threads = 256;
  size_t blocks = (n4 + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const double inv_count = 1.0 / static_cast<double>(rows);
  auto g4 = reinterpret_cast<const float4*>(g);
  auto m4 = reinterpret_cast<const float4*>(mask);
  auto y4 = reinterpret_cast<const float4*>(y);
  auto d4 = reinterpret_cast<float4*>(dy);
  if (round_tf32)
    bn_bwd_apply_kernel<true><<<(unsigned)blocks, threads, 0, stream>>>(
        g4, m4, y4, mean, invstd, gamma, sums, d4, dgamma, dbeta, n4, C, inv_count);
  else
    bn_bwd_apply_kernel<false><<<(unsigned)blocks, threads, 0, stream>>>(
        g4, m4, y4, mean, invstd, gamma, sums, d4, dgamma, dbeta, n4, C, inv_count);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("bn_bwd_apply: %s", cudaGetErrorString(e));
  return 0;
}

// ----------------------------------------------------------- zero-stuffing
// up[n, 2p, 2q, :] = dy[n, p, q, :], everything else 0  (turns a stride-2 data gradient into a
// stride-1 convolution over `up`).
__global__ void upsample_zero_kernel(const float4* __restrict__ dy, float4* __restrict__ up, int N,
                                     int P, int Q, int H, int W, int C4) {
  const size_t total = static_cast<size_t>(N) * H * W * C4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += stride) {
    const int c = static_cast<int>(i % C4);
    size_t t = i / C4;
    const int w = static_cast<int>(t % W); t /= W;
    const int h = static_cast<int>(t % H);
    const int n = static_cast<int>(t / H);
    float4 v = make_float4(0, 0, 0, 0);
    if (((h | w) & 1) == 0 && (h >> 1) < P && (w >> 1) < Q)
      v = dy[((static_cast<size_t>(n) * P + (h >> 1)) * Q + (w >> 1)) * C4 + c];
    up[i] = v;
  }
}

int launch_upsample_zero(const float* dy, float* up, int N, int P, int Q, int H, int W, int C,
                         cudaStream_t stream) {
  if (C % 4 != 0) return set_error("upsample_zero: C %% 4 != 0");
  const size_t total = static_cast<size_t>(N) * H * W * (C / 4);
  size_t blocks = (total + 255) / 256;
  const size_t cap = static_cast<size_t>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  upsample_zero_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
      reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(up), N, P, Q, H, W, C / 4);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("upsample_zero: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace b2n
