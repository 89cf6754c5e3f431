// Multi-tensor optimizer steps: one launch updates (up to 64 tensors of) a whole parameter list.
//
// SURVEY.md section 8f rank 1 ("next" row after the conv path): the reference steps
// torch.optim.Adam(lr 1e-4, wd 1e-4) in the consistency loops (eval_BreastPathQ_SSL_CR.py:481,100)
// and torch.optim.SGD(lr .01, momentum .9, nesterov, wd 1e-4) in the pretext loop
// (pretrain_BreastPathQ.py:245,61) -- in torch 1.7 a Python loop of ~6 tiny kernels per tensor
// over 62-66 tensors.  The arithmetic below follows torch.optim's single-tensor formulas term by
// term (same operation order, so results agree to the last bit or two); `grad_scale` folds the
// data-parallel 1/world averaging of the all-reduced gradient into the same pass.
#include <math.h>

#include "launch.h"

namespace b2n {

constexpr int kOptMaxTensors = 64;
struct OptTable {
  float* p[kOptMaxTensors];
  const float* g[kOptMaxTensors];
  float* s1[kOptMaxTensors];  // Adam exp_avg / SGD momentum buffer
  float* s2[kOptMaxTensors];  // Adam exp_avg_sq
  long long numel[kOptMaxTensors];
};

__global__ void counter_add_kernel(long long* c, long long d) { *c += d; }

// torch/optim/adam.py _single_tensor_adam (amsgrad=False, maximize=False, L2 weight decay).
// step_dev != null (CUDA-graph capturable mode): the step count lives on the device, so the bias
// corrections are formed here (in double, like the host path) instead of being baked into the launch.
__global__ void adam_multi_kernel(const OptTable t, float step_size, float one_minus_beta1, float beta2,
                                  float one_minus_beta2, float eps, float weight_decay,
                                  float bias_c2_sqrt, float grad_scale,
                                  const long long* __restrict__ step_dev, double lr, double beta1_d,
                                  double beta2_d) {
  if (step_dev != nullptr) {
    const double step = static_cast<double>(*step_dev);
    step_size = static_cast<float>(lr / (1.0 - pow(beta1_d, step)));
    bias_c2_sqrt = static_cast<float>(sqrt(1.0 - pow(beta2_d, step)));
  }
  float* __restrict__ p = t.p[blockIdx.y];
  const float* __restrict__ g = t.g[blockIdx.y];
  float* __restrict__ m = t.s1[blockIdx.y];
  float* __restrict__ v = t.s2[blockIdx.y];
  const long long n = t.numel[blockIdx.y];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const float pi = p[i];
    float gi = g[i] * grad_scale;
    if (weight_decay != 0.f) gi = fmaf(pi, weight_decay, gi);       // grad.add(param, alpha=wd)
    const float mi = fmaf(gi - m[i], one_minus_beta1, m[i]);        // exp_avg.lerp_(grad, 1-beta1)
    const float vi = fmaf(gi * gi, one_minus_beta2, v[i] * beta2);  // mul_(beta2).addcmul_(g, g, 1-beta2)
    const float denom = sqrtf(vi) / bias_c2_sqrt + eps;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / denom);                           // addcdiv_(exp_avg, denom, -step_size)
  }
}

// torch/optim/sgd.py _single_tensor_sgd (dampening 0, maximize=False)
__global__ void sgd_multi_kernel(const OptTable t, float lr, float momentum, float weight_decay,
                                 int nesterov, int first_step, float grad_scale) {
  float* __restrict__ p = t.p[blockIdx.y];
  const float* __restrict__ g = t.g[blockIdx.y];
  float* __restrict__ buf = t.s1[blockIdx.y];
  const long long n = t.numel[blockIdx.y];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const float pi = p[i];
    float gi = g[i] * grad_scale;
    if (weight_decay != 0.f) gi = fmaf(pi, weight_decay, gi);
    if (momentum != 0.f) {
      const float b = first_step ? gi : fmaf(buf[i], momentum, gi);  // buf.mul_(momentum).add_(grad)
      buf[i] = b;
      gi = nesterov ? fmaf(b, momentum, gi) : b;                     // grad.add(buf, alpha=momentum)
    }
    p[i] = fmaf(gi, -lr, pi);                                        // param.add_(grad, alpha=-lr)
  }
}

template <typename F>
static int for_each_chunk(float* const* p, const float* const* g, float* const* s1,
                          float* const* s2, const long long* numel, int n, const char* what, F&& f) {
  for (int base = 0; base < n; base += kOptMaxTensors) {
    OptTable t;
    const int cnt = n - base < kOptMaxTensors ? n - base : kOptMaxTensors;
    long long biggest = 1;
    for (int i = 0; i < kOptMaxTensors; ++i) {
      const int j = i < cnt ? base + i : base;
      t.p[i] = p[j];
      t.g[i] = g[j];
      t.s1[i] = s1 ? s1[j] : nullptr;
      t.s2[i] = s2 ? s2[j] : nullptr;
      t.numel[i] = i < cnt ? numel[j] : 0;
      if (t.numel[i] > biggest) biggest = t.numel[i];
    }
    long long bx = (biggest + 256 * 4 - 1) / (256 * 4);
    if (bx > 256) bx = 256;
    if (bx < 1) bx = 1;
    f(t, dim3(static_cast<unsigned>(bx), cnt));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("%s: %s", what, cudaGetErrorString(e));
  }
  return 0;
}

// Hyper-parameters arrive as the Python doubles torch.optim derives its scalars from: 1 - beta,
// lr / bias_correction1 and sqrt(bias_correction2) are formed in double and only then rounded.
int launch_adam_multi(float* const* p, const float* const* g, float* const* exp_avg,
                      float* const* exp_avg_sq, const long long* numel, int n, double lr, double beta1,
                      double beta2, double eps, double weight_decay, long long step,
                      long long* step_dev, double grad_scale, cudaStream_t stream) {
  float step_size = 0.f, bias_c2_sqrt = 1.f;
  if (step_dev != nullptr) {
    counter_add_kernel<<<1, 1, 0, stream>>>(step_dev, 1);   // this call is step *step_dev + 1
  } else {
    if (step < 1) return set_error("adam_multi: step must be >= 1 (got %lld)", step);
    const double bias_c1 = 1.0 - pow(beta1, static_cast<double>(step));
    const double bias_c2 = 1.0 - pow(beta2, static_cast<double>(step));
    step_size = static_cast<float>(lr / bias_c1);
    bias_c2_sqrt = static_cast<float>(sqrt(bias_c2));
  }
  return for_each_chunk(p, g, exp_avg, exp_avg_sq, numel, n, "adam_multi",
                        [&](const OptTable& t, dim3 grid) {
                          adam_multi_kernel<<<grid, 256, 0, stream>>>(
                              t, step_size, static_cast<float>(1.0 - beta1), static_cast<float>(beta2),
                              static_cast<float>(1.0 - beta2), static_cast<float>(eps),
                              static_cast<float>(weight_decay), bias_c2_sqrt,
                              static_cast<float>(grad_scale), step_dev, lr, beta1, beta2);
                        });
}

int launch_sgd_multi(float* const* p, const float* const* g, float* const* momentum_buf,
                     const long long* numel, int n, double lr, double momentum, double weight_decay,
                     int nesterov, int first_step, double grad_scale, cudaStream_t stream) {
  if (momentum != 0.0 && momentum_buf == nullptr)
    return set_error("sgd_multi: momentum needs momentum buffers");
  return for_each_chunk(p, g, momentum_buf, nullptr, numel, n, "sgd_multi",
                        [&](const OptTable& t, dim3 grid) {
                          sgd_multi_kernel<<<grid, 256, 0, stream>>>(
                              t, static_cast<float>(lr), static_cast<float>(momentum),
                              static_cast<float>(weight_decay), nesterov, first_step,
                              static_cast<float>(grad_scale));
                        });
}

}  // namespace b2n
