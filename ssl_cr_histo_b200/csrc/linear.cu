// FP32 SIMT GEMM for the small fully-connected heads: the pair-MLP Linear(1024,512)+ReLU+
// Linear(512,256) (models/net.py:36-37,60-62), Classifier Linear(768,128)+ReLU+Linear(128,C)
// (models/net.py:12-15) and FinetuneResNet Linear(768,C) (models/net.py:110).  These are
// < 0.1 % of the step's FLOPs, so they stay in exact FP32 FMA arithmetic (no TF32 rounding in
// front of the logits) -- 64x64x16 tiles, 4x4 outputs per thread.
#include "launch.h"

namespace b2n {

// C[m][n] = act(C_prev[m][n] (when accumulate) + sum_k A(m,k) * B(k,n) + bias[n]) * [mask[m][n] > 0]
// -- the previous content is added BEFORE bias / ReLU / mask, so a product split over K
// (cat(A,B) W^T = A Wa^T + B Wb^T) is two launches, the second carrying the bias and the ReLU.
// A(m,k) = a[m*sam + k*sak], B(k,n) = b[k*sbk + n*sbn], C row-major with pitch ldc.
struct GemmArgs {
  const float* a; long long sam, sak;
  const float* b; long long sbk, sbn;
  float* c; long long ldc;
  const float* bias;   // [N] or null
  const float* mask;   // [M][ldc] or null
  int M, N, K;
  int relu;
  int accumulate;
};

constexpr int TN = 64, TK = 16;

// RM = rows of the micro-tile (4: 64 x 64 block tile; 2: 32 x 64, for launches whose 64-row grid
// would leave most of the 148 SMs idle -- the heads' GEMMs have only 704 x 256..1024 outputs).
// The next K slab is fetched into registers while the current one is multiplied, and the
// micro-tile operands leave shared memory as one 128-bit (64-bit for RM = 2) load each: rows are
// padded by 4 floats, which keeps them 16-byte aligned and the slab stores at most 2-way conflicted.
// Every output is still one fmaf chain over k = 0..K-1, so results do not depend on the tiling.
template <int RM>
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs g) {
  constexpr int TM = 16 * RM;
  constexpr int NA = (TM * TK) / 256, NB = (TN * TK) / 256;
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each an RM x 4 micro-tile
  float acc[RM][4] = {};
  const bool a_kfast = g.sak == 1;  // choose the smem fill order that keeps gmem reads coalesced
  const bool b_kfast = g.sbk == 1;
  float ra[NA], rb[NB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int e = tid + i * 256;
      const int mm = a_kfast ? e / TK : e % TM;
      const int kk = a_kfast ? e % TK : e / TM;
      const int m = m0 + mm, k = k0 + kk;
      ra[i] = (m < g.M && k < g.K) ? __ldg(g.a + m * g.sam + k * g.sak) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = tid + i * 256;
      const int nn = b_kfast ? e / TK : e % TN;
      const int kk = b_kfast ? e % TK : e / TN;
      const int n = n0 + nn, k = k0 + kk;
      rb[i] = (n < g.N && k < g.K) ? __ldg(g.b + k * g.sbk + n * g.sbn) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < g.K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int e = tid + i * 256;
      As[a_kfast ? e % TK : e / TM][a_kfast ? e / TK : e % TM] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = tid + i * 256;
      Bs[b_kfast ? e % TK : e / TN][b_kfast ? e / TK : e % TN] = rb[i];
    }
    __syncthreads();
    if (k0 + TK < g.K) fetch(k0 + TK);
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[RM], bv[4];
      if constexpr (RM == 4) {
        const float4 t = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
      } else if constexpr (RM == 2) {
        const float2 t = *reinterpret_cast<const float2*>(&As[kk][ty * 2]);
        av[0] = t.x; av[1] = t.y;
      } else {
#pragma unroll
        for (int i = 0; i < RM; ++i) av[i] = As[kk][ty * RM + i];
      }
      {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int m = m0 + ty * RM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      const size_t o = static_cast<size_t>(m) * g.ldc + n;
      if (g.accumulate) v += g.c[o];
      if (g.bias != nullptr) v += g.bias[n];
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.mask != nullptr && !(g.mask[o] > 0.f)) v = 0.f;
      g.c[o] = v;
    }
  }
}

static int run_gemm(const GemmArgs& g, cudaStream_t stream, const char* who) {
  if (g.M <= 0 || g.N <= 0) return 0;
  const long long blocks64 = 1ll * ((g.N + TN - 1) / TN) * ((g.M + 63) / 64);
  if (blocks64 < 2ll * device_sm_count()) {
    dim3 grid((g.N + TN - 1) / TN, (g.M + 31) / 32);
    sgemm_kernel<2><<<grid, 256, 0, stream>>>(g);
  } else {
    dim3 grid((g.N + TN - 1) / TN, (g.M + 63) / 64);
    sgemm_kernel<4><<<grid, 256, 0, stream>>>(g);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("%s: %s", who, cudaGetErrorString(e));
  return 0;
}

// y[n][o] = act(x[n][:] . w[o][:] + b[o]);  x pitch ldx, y pitch ldy, w pitch ldw
int launch_linear_fwd(const float* x, long long ldx, const float* w, long long ldw, const float* b,
                      float* y, long long ldy, int rows, int in_f, int out_f, int relu,
                      int accumulate, cudaStream_t stream) {
  GemmArgs g{x, ldx, 1, w, 1, ldw, y, ldy, b, nullptr, rows, out_f, in_f, relu, accumulate};
  return run_gemm(g, stream, "linear_fwd");
}
// dx[n][i] = sum_o dy[n][o] * w[o][i]   (optionally gated by mask > 0, optionally accumulated)
int launch_linear_bwd_data(const float* dy, long long lddy, const float* w, long long ldw,
                           float* dx, long long lddx, const float* mask, int rows, int in_f,
                           int out_f, int accumulate, cudaStream_t stream) {
  GemmArgs g{dy, lddy, 1, w, ldw, 1, dx, lddx, nullptr, mask, rows, in_f, out_f, 0, accumulate};
  return run_gemm(g, stream, "linear_bwd_data");
}
// dw[o][i] (+)= sum_n dy[n][o] * x[n][i]
int launch_linear_bwd_weight(const float* dy, long long lddy, const float* x, long long ldx,
                             float* dw, long long lddw, int rows, int in_f, int out_f,
                             int accumulate, cudaStream_t stream) {
  GemmArgs g{dy, 1, lddy, x, ldx, 1, dw, lddw, nullptr, nullptr, out_f, in_f, rows, 0, accumulate};
  return run_gemm(g, stream, "linear_bwd_weight");
}

// Column-block helpers of the pair head when its three pair rows are identical
// (TripletNet_Finetune, models/net.py:92-103): the (rows, copies*width) feature matrix holds `copies`
// equal blocks side by side -- filled from block 0 -- and its gradient is the sum of the blocks.
__global__ void cols_replicate_kernel(float* __restrict__ y, long long ld, int rows, int width,
                                      int copies) {
  const long long total = static_cast<long long>(rows) * width;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = t / width;
    const int c = static_cast<int>(t - r * width);
    const float v = y[r * ld + c];
    for (int k = 1; k < copies; ++k) y[r * ld + k * width + c] = v;
  }
}
__global__ void cols_sum_kernel(const float* __restrict__ dy, long long ld, float* __restrict__ out,
                                int rows, int width, int copies) {
  const long long total = static_cast<long long>(rows) * width;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = t / width;
    const int c = static_cast<int>(t - r * width);
    float v = dy[r * ld + c];
    for (int k = 1; k < copies; ++k) v += dy[r * ld + k * width + c];   // left to right, like autograd
    out[t] = v;
  }
}
static unsigned cols_grid(long long total) {
  long long b = (total + 255) / 256;
  if (b > 1184) b = 1184;
  return static_cast<unsigned>(b < 1 ? 1 : b);
}
int launch_cols_replicate(float* y, long long ld, int rows, int width, int copies, cudaStream_t stream) {
  if (rows <= 0 || width <= 0 || copies < 1) return rows == 0 ? 0 : set_error("cols_replicate: bad shape");
  cols_replicate_kernel<<<cols_grid(1ll * rows * width), 256, 0, stream>>>(y, ld, rows, width, copies);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("cols_replicate: %s", cudaGetErrorString(e));
  return 0;
}
int launch_cols_sum(const float* dy, long long ld, float* out, int rows, int width, int copies,
                    cudaStream_t stream) {
  if (rows <= 0 || width <= 0 || copies < 1) return rows == 0 ? 0 : set_error("cols_sum: bad shape");
  cols_sum_kernel<<<cols_grid(1ll * rows * width), 256, 0, stream>>>(dy, ld, out, rows, width, copies);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("cols_sum: %s", cudaGetErrorString(e));
  return 0;
}

// db[o] (+)= sum_n dy[n][o].  A block owns 32 columns; its 16 row slices (rows ry, ry + 16, ...)
// are summed by 16 threads per column and combined in a fixed order: deterministic, and 16 loads
// in flight per column instead of one serial chain over all rows.
constexpr int kColsumSlices = 16;
__global__ void __launch_bounds__(32 * kColsumSlices)
colsum_kernel(const float* __restrict__ dy, long long lddy, float* __restrict__ db, int rows, int out_f,
              int accumulate) {
  __shared__ float part[kColsumSlices][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + cx;
  float acc = 0.f;
  if (o < out_f)
    for (int n = ry; n < rows; n += kColsumSlices) acc += dy[static_cast<size_t>(n) * lddy + o];
  part[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && o < out_f) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kColsumSlices; ++i) t += part[i][cx];
    db[o] = accumulate ? db[o] + t : t;
  }
}
int launch_colsum(const float* dy, long long lddy, float* db, int rows, int out_f, int accumulate,
                  cudaStream_t stream) {
  colsum_kernel<<<(out_f + 31) / 32, 32 * kColsumSlices, 0, stream>>>(dy, lddy, db, rows, out_f, accumulate);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("colsum: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace b2n
