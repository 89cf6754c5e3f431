// BatchNorm + ReLU + residual element-wise kernels (NHWC fp32, HBM-bound, 128-bit accesses).
//
// They replace cudnnBatchNormalizationForwardTraining/Backward and the ATen ReLU / add kernels
// behind torchvision BasicBlock.forward (site-packages/torchvision/models/resnet.py:93-103).
// The per-channel batch statistics themselves (sum, sum of squares) come out of the conv
// kernel's epilogue (conv_igemm.cuh); here they are finalised, applied, and differentiated.
#include <cuda_fp16.h>

#include "launch.h"
#include "ptx.cuh"

namespace b2n {

// ------------------------------------------------------------------ finalize
// stats = [2][C] doubles (sum, sumsq) over `count` values per channel.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, float* __restrict__ inv_gamma,
                                   int C, double count, float momentum, float eps, int n_updates) {
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
  // 1 / gamma: lets a consumer that only sees the activation a = gamma * xhat + beta recover xhat
  // (the stem's BatchNorm-backward sums are taken over the max-pooled activation)
  if (inv_gamma != nullptr) inv_gamma[c] = gamma[c] != 0.f ? 1.f / gamma[c] : 0.f;
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

// The same for every BatchNorm layer of a model in one launch (an eval-mode trunk pass folds all
// twenty): block b of the grid serves 128 channels of the layer that owns it.
constexpr int kMaxFoldJobs = 32;
struct FoldJob {
  const float *gamma, *beta, *rm, *rv;
  float *scale, *shift;
  int C;
  float eps;
  int block0;
};
struct FoldTable {
  FoldJob job[kMaxFoldJobs];
  int n;
};
__global__ void __launch_bounds__(128) bn_fold_eval_multi_kernel(const __grid_constant__ FoldTable tbl) {
  int j = 0;
  while (j + 1 < tbl.n && static_cast<int>(blockIdx.x) >= tbl.job[j + 1].block0) ++j;
  const FoldJob& q = tbl.job[j];
  const int c = (blockIdx.x - q.block0) * 128 + threadIdx.x;
  if (c >= q.C) return;
  const float sc = q.gamma[c] / sqrtf(q.rv[c] + q.eps);
  q.scale[c] = sc;
  q.shift[c] = q.beta[c] - q.rm[c] * sc;
}

// --------------------------------------------------------------------- apply
// v = act(scale*y + shift + residual); residual = res32 (identity / raw tensor), or
// res_scale*res32 + res_shift (the downsample branch's BN), or the (hi, lo) FP16 pair
// res_h + res_l (identity shortcut of a forward activation).
// Outputs (each optional): out32 -- fp32, TF32-rounded when ROUND (the backward pass' operand /
// ReLU mask); out_h, out_l -- the (hi, lo) FP16 pair the next forward conv consumes.
// 8 channels per thread: 2 x 128-bit fp32 loads, 1 x 128-bit store per FP16 plane.
// Two 8-channel groups per thread and iteration, all loads issued before the first use (the kernel
// is pure streaming: memory-level parallelism is what sets its bandwidth); y and the fp32 copy are
// touched once per pass, so they use the streaming (evict-first) cache hints and leave L2 to the
// FP16 pair the next conv reads.
template <bool RELU, bool ROUND>
__global__ void bn_apply_kernel(const float4* __restrict__ y, const float* __restrict__ scale,
                                const float* __restrict__ shift, const float4* __restrict__ res32,
                                const float* __restrict__ res_scale,
                                const float* __restrict__ res_shift, const uint4* __restrict__ res_h,
                                const uint4* __restrict__ res_l, float4* __restrict__ out32,
                                uint4* __restrict__ out_h, uint4* __restrict__ out_l, size_t n8,
                                int C) {
  constexpr int U = 2;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i0 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < n8; i0 += U * stride) {
    float4 y0[U], y1[U], r0[U], r1[U];
    uint4 rh[U], rl[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * stride;
      ok[u] = i < n8;
      if (!ok[u]) continue;
      y0[u] = __ldcs(y + 2 * i);
      y1[u] = __ldcs(y + 2 * i + 1);
      if (res32 != nullptr) { r0[u] = __ldcs(res32 + 2 * i); r1[u] = __ldcs(res32 + 2 * i + 1); }
      if (res_h != nullptr) { rh[u] = res_h[i]; rl[u] = res_l[i]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      const size_t i = i0 + u * stride;
      const int c = static_cast<int>((i * 8) % C);
      const float4 sc0 = *reinterpret_cast<const float4*>(scale + c);
      const float4 sc1 = *reinterpret_cast<const float4*>(scale + c + 4);
      const float4 sh0 = *reinterpret_cast<const float4*>(shift + c);
      const float4 sh1 = *reinterpret_cast<const float4*>(shift + c + 4);
      float o[8] = {fmaf(y0[u].x, sc0.x, sh0.x), fmaf(y0[u].y, sc0.y, sh0.y), fmaf(y0[u].z, sc0.z, sh0.z),
                    fmaf(y0[u].w, sc0.w, sh0.w), fmaf(y1[u].x, sc1.x, sh1.x), fmaf(y1[u].y, sc1.y, sh1.y),
                    fmaf(y1[u].z, sc1.z, sh1.z), fmaf(y1[u].w, sc1.w, sh1.w)};
      if (res32 != nullptr) {
        float r[8] = {r0[u].x, r0[u].y, r0[u].z, r0[u].w, r1[u].x, r1[u].y, r1[u].z, r1[u].w};
        if (res_scale != nullptr) {
#pragma unroll
          for (int k = 0; k < 8; ++k) r[k] = fmaf(r[k], res_scale[c + k], res_shift[c + k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += r[k];
      }
      if (res_h != nullptr) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&rh[u]);
        const __half2* l2 = reinterpret_cast<const __half2*>(&rl[u]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = __half22float2(h2[k]), b = __half22float2(l2[k]);
          o[2 * k] += a.x + b.x;
          o[2 * k + 1] += a.y + b.y;
        }
      }
      if (RELU) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaxf(o[k], 0.f);
      }
      if (out_h != nullptr) {
        uint4 ph, pl;
        __half2* h2 = reinterpret_cast<__half2*>(&ph);
        __half2* l2 = reinterpret_cast<__half2*>(&pl);
#pragma unroll
        for (int k = 0; k < 4; ++k) split_f16(o[2 * k], o[2 * k + 1], h2[k], l2[k]);
        out_h[i] = ph;
        out_l[i] = pl;
      }
      if (out32 != nullptr) {
        if (ROUND) {
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = tf32_rn(o[k]);
        }
        __stcs(out32 + 2 * i, make_float4(o[0], o[1], o[2], o[3]));
        __stcs(out32 + 2 * i + 1, make_float4(o[4], o[5], o[6], o[7]));
      }
    }
  }
}

int launch_bn_apply(const float* y, const float* scale, const float* shift, const float* res32,
                    const float* res_scale, const float* res_shift, const __half* res_h,
                    const __half* res_l, float* out32, __half* out_h, __half* out_l,
                    long long rows, int C, int relu, int round_tf32, cudaStream_t stream) {
  if (C % 8 != 0) return set_error("bn_apply: C %% 8 != 0");
  if ((out_h != nullptr) != (out_l != nullptr) || (res_h != nullptr) != (res_l != nullptr))
    return set_error("bn_apply: FP16 pairs need both planes");
  const size_t n8 = static_cast<size_t>(rows) * C / 8;
  if (n8 == 0) return 0;
  const int threads = 256;
  size_t blocks = (n8 + 2 * threads - 1) / (2 * threads);   // two groups per thread and iteration
  const size_t cap = static_cast<size_t>(device_sm_count()) * elementwise_blocks_per_sm();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
#define B2N_LAUNCH(R, T)                                                                        \
  bn_apply_kernel<R, T><<<(unsigned)blocks, threads, 0, stream>>>(                              \
      reinterpret_cast<const float4*>(y), scale, shift, reinterpret_cast<const float4*>(res32), \
      res_scale, res_shift, reinterpret_cast<const uint4*>(res_h),                              \
      reinterpret_cast<const uint4*>(res_l), reinterpret_cast<float4*>(out32),                  \
      reinterpret_cast<uint4*>(out_h), reinterpret_cast<uint4*>(out_l), n8, C)
  if (relu && round_tf32) B2N_LAUNCH(true, true);
  else if (relu) B2N_LAUNCH(true, false);
  else if (round_tf32) B2N_LAUNCH(false, true);
  else B2N_LAUNCH(false, false);
#undef B2N_LAUNCH
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("bn_apply: %s", cudaGetErrorString(e));
  return 0;
}

int launch_bn_finalize(const double* stats, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float* scale, float* shift,
                       float* mean, float* invstd, float* inv_gamma, int C, double count,
                       float momentum, float eps, int n_updates, cudaStream_t stream) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(stats, gamma, beta, running_mean,
                                                         running_var, scale, shift, mean, invstd,
                                                         inv_gamma, C, count, momentum, eps, n_updates);
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

int launch_bn_fold_eval_multi(const float* const* gamma, const float* const* beta, const float* const* rm,
                              const float* const* rv, float* const* scale, float* const* shift,
                              const int* C, const float* eps, int n, cudaStream_t stream) {
  for (int base = 0; base < n; base += kMaxFoldJobs) {
    FoldTable tbl;
    tbl.n = n - base < kMaxFoldJobs ? n - base : kMaxFoldJobs;
    int blocks = 0;
    for (int i = 0; i < tbl.n; ++i) {
      const int g = base + i;
      if (!gamma[g] || !beta[g] || !rm[g] || !rv[g] || !scale[g] || !shift[g] || C[g] <= 0)
        return set_error("bn_fold_eval_multi: bad job %d", g);
      FoldJob& q = tbl.job[i];
      q.gamma = gamma[g]; q.beta = beta[g]; q.rm = rm[g]; q.rv = rv[g];
      q.scale = scale[g]; q.shift = shift[g];
      q.C = C[g]; q.eps = eps[g]; q.block0 = blocks;
      blocks += (C[g] + 127) / 128;
    }
    bn_fold_eval_multi_kernel<<<blocks, 128, 0, stream>>>(tbl);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("bn_fold_eval_multi: %s", cudaGetErrorString(e));
  }
  return 0;
}

// ------------------------------------------------------------------ backward
// g' = g * (mask > 0)   (mask = the block's post-ReLU output; null = no ReLU gate).  When the
// ReLU input is the BN output itself (no residual: bn1 of a BasicBlock) the gate can instead be
// recomputed from y -- fmaf(y, gate_scale, gate_shift) > 0, the forward's own expression -- which
// saves reading the mask tensor in both passes.
// sums[0][c] = sum g',  sums[1][c] = sum g' * xhat,  xhat = (y - mean) * invstd
//
// Thread layout: threadIdx.x -> float4 channel group (C/4 of them), threadIdx.y -> row lane.
__global__ void bn_bwd_reduce_kernel(const float4* __restrict__ g, const float4* __restrict__ mask,
                                     const float4* __restrict__ y, const float* __restrict__ mean,
                                     const float* __restrict__ invstd,
                                     const float* __restrict__ gate_scale,
                                     const float* __restrict__ gate_shift, double* __restrict__ sums,
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
    float4 gs = make_float4(0, 0, 0, 0), gh = gs;
    if (gate_scale != nullptr) {
      gs = *reinterpret_cast<const float4*>(gate_scale + 4 * cg);
      gh = *reinterpret_cast<const float4*>(gate_shift + 4 * cg);
    }
    for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      const size_t i = static_cast<size_t>(r) * C4 + cg;
      float4 gv = g[i];
      if (mask != nullptr) {
        const float4 m = mask[i];
        gv.x = m.x > 0.f ? gv.x : 0.f; gv.y = m.y > 0.f ? gv.y : 0.f;
        gv.z = m.z > 0.f ? gv.z : 0.f; gv.w = m.w > 0.f ? gv.w : 0.f;
      }
      const float4 yv = y[i];
      if (gate_scale != nullptr) {
        gv.x = fmaf(yv.x, gs.x, gh.x) > 0.f ? gv.x : 0.f; gv.y = fmaf(yv.y, gs.y, gh.y) > 0.f ? gv.y : 0.f;
        gv.z = fmaf(yv.z, gs.z, gh.z) > 0.f ? gv.z : 0.f; gv.w = fmaf(yv.w, gs.w, gh.w) > 0.f ? gv.w : 0.f;
      }
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
// The per-channel constants (gamma*invstd, sum_g/M, sum_gx/M from the FP64 sums, mean, invstd, the
// optional gate affine) are formed once per block in shared memory -- not per element.
template <bool ROUND>
__global__ void bn_bwd_apply_kernel(const float4* __restrict__ g, const float4* __restrict__ mask,
                                    const float4* __restrict__ y, const float* __restrict__ mean,
                                    const float* __restrict__ invstd,
                                    const float* __restrict__ gamma,
                                    const float* __restrict__ gate_scale,
                                    const float* __restrict__ gate_shift,
                                    const double* __restrict__ sums, float4* __restrict__ dy,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, size_t n4,
                                    int C, double inv_count, int accumulate) {
  extern __shared__ float chan[];  // [7][C]: gi, m1, m2, mean, invstd, gate scale, gate shift
  float* s_gi = chan;
  float* s_m1 = chan + C;
  float* s_m2 = chan + 2 * C;
  float* s_mu = chan + 3 * C;
  float* s_is = chan + 4 * C;
  float* s_gs = chan + 5 * C;
  float* s_gh = chan + 6 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float is = invstd[c];
    s_gi[c] = gamma[c] * is;
    s_m1[c] = static_cast<float>(sums[c] * inv_count);
    s_m2[c] = static_cast<float>(sums[C + c] * inv_count);
    s_mu[c] = mean[c];
    s_is[c] = is;
    s_gs[c] = gate_scale != nullptr ? gate_scale[c] : 0.f;
    s_gh[c] = gate_shift != nullptr ? gate_shift[c] : 0.f;
    if (blockIdx.x == 0 && dgamma != nullptr) {
      // accumulate: += into a gradient slot that other passes over the same weights also feed
      dbeta[c] = (accumulate ? dbeta[c] : 0.f) + static_cast<float>(sums[c]);
      dgamma[c] = (accumulate ? dgamma[c] : 0.f) + static_cast<float>(sums[C + c]);
    }
  }
  __syncthreads();
  const bool gated = gate_scale != nullptr;
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
    if (gated) {
      const float4 gs = *reinterpret_cast<const float4*>(s_gs + c);
      const float4 gh = *reinterpret_cast<const float4*>(s_gh + c);
      gv.x = fmaf(yv.x, gs.x, gh.x) > 0.f ? gv.x : 0.f; gv.y = fmaf(yv.y, gs.y, gh.y) > 0.f ? gv.y : 0.f;
      gv.z = fmaf(yv.z, gs.z, gh.z) > 0.f ? gv.z : 0.f; gv.w = fmaf(yv.w, gs.w, gh.w) > 0.f ? gv.w : 0.f;
    }
    const float4 mu = *reinterpret_cast<const float4*>(s_mu + c);
    const float4 is = *reinterpret_cast<const float4*>(s_is + c);
    const float4 gi = *reinterpret_cast<const float4*>(s_gi + c);
    const float4 m1 = *reinterpret_cast<const float4*>(s_m1 + c);
    const float4 m2 = *reinterpret_cast<const float4*>(s_m2 + c);
    float o[4];
    const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
    const float mm[4] = {mu.x, mu.y, mu.z, mu.w};
    const float ii[4] = {is.x, is.y, is.z, is.w};
    const float aa[4] = {gi.x, gi.y, gi.z, gi.w};
    const float a1[4] = {m1.x, m1.y, m1.z, m1.w};
    const float a2[4] = {m2.x, m2.y, m2.z, m2.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (yy[k] - mm[k]) * ii[k];
      const float v = aa[k] * (gg[k] - a1[k] - xh * a2[k]);
      o[k] = ROUND ? tf32_rn(v) : v;
    }
    dy[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

int launch_bn_bwd_reduce(const float* g, const float* mask, const float* y, const float* mean,
                         const float* invstd, const float* gate_scale, const float* gate_shift,
                         double* sums, long long rows, int C, cudaStream_t stream) {
  if ((gate_scale != nullptr) != (gate_shift != nullptr))
    return set_error("bn_bwd_reduce: gate_scale and gate_shift come as a pair");
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
      reinterpret_cast<const float4*>(y), mean, invstd, gate_scale, gate_shift, sums, rows, C,
      static_cast<int>(rpb));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("bn_bwd_reduce: %s", cudaGetErrorString(e));
  return 0;
}

int launch_bn_bwd_apply(const float* g, const float* mask, const float* y, const float* mean,
                        const float* invstd, const float* gamma, const float* gate_scale,
                        const float* gate_shift, const double* sums, float* dy, float* dgamma,
                        float* dbeta, long long rows, int C, int round_tf32, int accumulate,
                        cudaStream_t stream) {
  if ((gate_scale != nullptr) != (gate_shift != nullptr))
    return set_error("bn_bwd_apply: gate_scale and gate_shift come as a pair");
  if (C % 4 != 0) return set_error("bn_bwd_apply: C %% 4 != 0");
  const size_t n4 = static_cast<size_t>(rows) * C / 4;
  const int threads = 256;
  size_t blocks = (n4 + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(device_sm_count()) * elementwise_blocks_per_sm();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const double inv_count = 1.0 / static_cast<double>(rows);
  auto g4 = reinterpret_cast<const float4*>(g);
  auto m4 = reinterpret_cast<const float4*>(mask);
  auto y4 = reinterpret_cast<const float4*>(y);
  auto d4 = reinterpret_cast<float4*>(dy);
  const size_t smem = static_cast<size_t>(7) * C * sizeof(float);
  if (smem > 48 * 1024) return set_error("bn_bwd_apply: C=%d too large", C);
  if (round_tf32)
    bn_bwd_apply_kernel<true><<<(unsigned)blocks, threads, smem, stream>>>(
        g4, m4, y4, mean, invstd, gamma, gate_scale, gate_shift, sums, d4, dgamma, dbeta, n4, C,
        inv_count, accumulate);
  else
    bn_bwd_apply_kernel<false><<<(unsigned)blocks, threads, smem, stream>>>(
        g4, m4, y4, mean, invstd, gamma, gate_scale, gate_shift, sums, d4, dgamma, dbeta, n4, C,
        inv_count, accumulate);
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
  const size_t cap = static_cast<size_t>(device_sm_count()) * elementwise_blocks_per_sm();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  upsample_zero_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
      reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(up), N, P, Q, H, W, C / 4);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("upsample_zero: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace b2n
