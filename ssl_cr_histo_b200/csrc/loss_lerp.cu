// Fused loss kernels and the multi-tensor weight lerp.
//
// Losses (one launch each, forward value + d(loss)/d(logits) + decisions):
//   * softmax cross-entropy, mean reduction, argmax      -- nn.CrossEntropyLoss + torch.argmax
//     (pretrain_BreastPathQ.py:56,66, eval_Kather_SSL.py:67,78)
//   * consistency step, MSE/MSE                          -- eval_BreastPathQ_SSL_CR.py:92-95
//   * consistency step, CE + CE-to-teacher-argmax        -- eval_Kather_SSL_CR.py:87-93
// Lerp: dst <- alpha*src + (1-alpha)*dst over a whole parameter list in one launch
//   * alpha = 1  : teacher <- student hand-off (copy.deepcopy, eval_BreastPathQ_SSL_CR.py:515)
//   * alpha = 1-la_alpha, write_back: Lookahead slow-weight pull (lookahead.py:96-97)
#include <math.h>

#include "launch.h"

namespace b2n {

constexpr int kMaxClasses = 32;

__device__ inline float block_sum(float v, float* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += scratch[w];
  return t;
}

// CE of one row; optionally writes grad = (softmax - onehot) * gscale and returns argmax.
// A target outside [0, C) (F.cross_entropy's ignore_index = -100 included: the reference never
// uses it) is not dereferenced; the row's loss and gradient become NaN so the step fails loudly
// instead of training on garbage.
__device__ inline float ce_row(const float* __restrict__ lg, int C, long long target,
                               float* __restrict__ grad, float gscale, int* amax) {
  float mx = lg[0];
  int am = 0;
  for (int c = 1; c < C; ++c)
    if (lg[c] > mx) { mx = lg[c]; am = c; }
  float se = 0.f;
  for (int c = 0; c < C; ++c) se += expf(lg[c] - mx);
  const float lse = logf(se) + mx;
  const bool ok = target >= 0 && target < C;
  if (grad != nullptr)
    for (int c = 0; c < C; ++c)
      grad[c] = ok ? (expf(lg[c] - lse) - (c == target ? 1.f : 0.f)) * gscale : __int_as_float(0x7fc00000);
  if (amax != nullptr) *amax = am;
  return ok ? lse - lg[target] : __int_as_float(0x7fc00000);
}

// mode 0: plain CE.  rows_x rows of logits_x vs targets (int64).
//   losses[0] = mean CE; dlogits_x; argmax_x
// mode 1: MSE consistency.  logits_x [rows_x][C] vs targets_f [rows_x*C]; logits_u_s vs logits_u_w.
//   losses = {sup, cons, sup + lambda*cons}; dlogits_x, dlogits_u (w.r.t. logits_u_s)
// mode 2: CE consistency.  sup = CE(logits_x, targets_i); pseudo = argmax(logits_u_w);
//   cons = CE(logits_u_s, pseudo); losses as mode 1; argmax_x, pseudo_out.
__global__ void __launch_bounds__(256)
fused_loss_kernel(int mode, const float* __restrict__ logits_x, const long long* __restrict__ targets_i,
                  const float* __restrict__ targets_f, const float* __restrict__ logits_u_w,
                  const float* __restrict__ logits_u_s, int rows_x, int rows_u, int C, float lambda_u,
                  float* __restrict__ losses, float* __restrict__ dlogits_x,
                  float* __restrict__ dlogits_u, long long* __restrict__ argmax_x,
                  long long* __restrict__ pseudo_out) {
  __shared__ float scratch[8];
  float sup = 0.f, cons = 0.f;
  if (mode == 1) {
    const int nx = rows_x * C, nu = rows_u * C;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      const float d = logits_x[i] - targets_f[i];
      sup += d * d;
      if (dlogits_x != nullptr) dlogits_x[i] = 2.f * d / static_cast<float>(nx);
    }
    for (int i = threadIdx.x; i < nu; i += blockDim.x) {
      const float d = logits_u_s[i] - logits_u_w[i];
      cons += d * d;
      if (dlogits_u != nullptr) dlogits_u[i] = 2.f * lambda_u * d / static_cast<float>(nu);
    }
    sup = block_sum(sup, scratch) / static_cast<float>(nx > 0 ? nx : 1);
    cons = block_sum(cons, scratch) / static_cast<float>(nu > 0 ? nu : 1);
  } else {
    for (int r = threadIdx.x; r < rows_x; r += blockDim.x) {
      int am;
      sup += ce_row(logits_x + static_cast<size_t>(r) * C, C, targets_i[r],
                    dlogits_x ? dlogits_x + static_cast<size_t>(r) * C : nullptr,
                    1.f / static_cast<float>(rows_x), &am);
      if (argmax_x != nullptr) argmax_x[r] = am;
    }
    sup = block_sum(sup, scratch) / static_cast<float>(rows_x > 0 ? rows_x : 1);
    if (mode == 2) {
      for (int r = threadIdx.x; r < rows_u; r += blockDim.x) {
        const float* lw = logits_u_w + static_cast<size_t>(r) * C;
        int pl = 0;
        for (int c = 1; c < C; ++c)
          if (lw[c] > lw[pl]) pl = c;  // argmax softmax == argmax logits (first maximum)
        if (pseudo_out != nullptr) pseudo_out[r] = pl;
        cons += ce_row(logits_u_s + static_cast<size_t>(r) * C, C, pl,
                       dlogits_u ? dlogits_u + static_cast<size_t>(r) * C : nullptr,
                       lambda_u / static_cast<float>(rows_u), nullptr);
      }
      cons = block_sum(cons, scratch) / static_cast<float>(rows_u > 0 ? rows_u : 1);
    }
  }
  if (threadIdx.x == 0) {
    losses[0] = sup;
    losses[1] = cons;
    losses[2] = sup + lambda_u * cons;
  }
}

int launch_fused_loss(int mode, const float* logits_x, const long long* targets_i,
                      const float* targets_f, const float* logits_u_w, const float* logits_u_s,
                      int rows_x, int rows_u, int C, float lambda_u, float* losses,
                      float* dlogits_x, float* dlogits_u, long long* argmax_x,
                      long long* pseudo_out, cudaStream_t stream) {
  if (mode < 0 || mode > 2) return set_error("fused_loss: bad mode %d", mode);
  if (C < 1 || C > kMaxClasses) return set_error("fused_loss: C=%d out of range", C);
  if (mode == 0) { rows_u = 0; lambda_u = 0.f; }
  fused_loss_kernel<<<1, 256, 0, stream>>>(mode, logits_x, targets_i, targets_f, logits_u_w,
                                           logits_u_s, rows_x, rows_u, C, lambda_u, losses,
                                           dlogits_x, dlogits_u, argmax_x, pseudo_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("fused_loss: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------- softmax probability of the last class
// out[i] = softmax(logits[i, :])[C-1]  -- the 'tumor' column of test_Camelyon16.py:57-58.
__global__ void softmax_last_kernel(const float* __restrict__ logits, float* __restrict__ out,
                                    int rows, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* l = logits + static_cast<size_t>(i) * C;
  float mx = l[0];
  for (int c = 1; c < C; ++c) mx = fmaxf(mx, l[c]);
  float den = 0.f;
  for (int c = 0; c < C; ++c) den += expf(l[c] - mx);
  out[i] = expf(l[C - 1] - mx) / den;
}

int launch_softmax_last(const float* logits, float* out, int rows, int C, cudaStream_t stream) {
  if (rows < 0 || C < 1) return set_error("softmax_last: bad shape (%d, %d)", rows, C);
  if (rows == 0) return 0;
  softmax_last_kernel<<<(rows + 127) / 128, 128, 0, stream>>>(logits, out, rows, C);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("softmax_last: %s", cudaGetErrorString(e));
  return 0;
}

// ---------------------------------------------------------------- multi lerp
constexpr int kLerpMaxTensors = 96;
struct LerpTable {
  float* dst[kLerpMaxTensors];
  float* src[kLerpMaxTensors];
  long long numel[kLerpMaxTensors];
};

__global__ void lerp_multi_kernel(const LerpTable t, float alpha, int write_back) {
  float* __restrict__ d = t.dst[blockIdx.y];
  float* __restrict__ s = t.src[blockIdx.y];
  const long long n = t.numel[blockIdx.y];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const float sv = s[i];
    const float v = alpha == 1.f ? sv : fmaf(alpha, sv, (1.f - alpha) * d[i]);
    d[i] = v;
    if (write_back) s[i] = v;
  }
}

int launch_lerp_multi(float* const* dst, float* const* src, const long long* numel, int n,
                      float alpha, int write_back, cudaStream_t stream) {
  for (int base = 0; base < n; base += kLerpMaxTensors) {
    LerpTable t;
    int cnt = n - base < kLerpMaxTensors ? n - base : kLerpMaxTensors;
    long long biggest = 1;
    for (int i = 0; i < kLerpMaxTensors; ++i) {
      const int j = i < cnt ? base + i : base;
      t.dst[i] = dst[j];
      t.src[i] = src[j];
      t.numel[i] = i < cnt ? numel[j] : 0;
      if (t.numel[i] > biggest) biggest = t.numel[i];
    }
    long long bx = (biggest + 256 * 4 - 1) / (256 * 4);
    if (bx > 512) bx = 512;
    if (bx < 1) bx = 1;
    dim3 grid(static_cast<unsigned>(bx), cnt);
    lerp_multi_kernel<<<grid, 256, 0, stream>>>(t, alpha, write_back);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("lerp_multi: %s", cudaGetErrorString(e));
  }
  return 0;
}

}  // namespace b2n
