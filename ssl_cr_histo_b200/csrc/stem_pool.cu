// Stem re-layout, max-pool and average-pool kernels (HBM-bound).
//
// The 7x7/stride-2 stem convolution (torchvision resnet.py:197, applied :268) is run on the
// tensor cores as a 4x4/stride-1 convolution over a 2x2 space-to-depth view of the input:
//   xs[n, i, j, (dy*2+dx)*3 + c] = x[n, c, 2i+dy, 2j+dx]         (12 real channels, zero padded)
//   ws[k, tr, ts, (dy*2+dx)*3 + c] = w[k, c, 2tr+dy-1, 2ts+dx-1] (zero when out of the 7x7 window)
//   out[p, q] = sum xs[p-2+tr, q-2+ts, :] . ws[:, tr, ts, :]      (padding 2 low / 1 high)
// The forward operand (hi, lo) FP16 pair is padded to 16 channels (32-byte pixel rows: one K = 16
// MMA per filter tap, four taps sharing one HALO box); the fp32 copy the weight gradient reads is
// padded to 32 channels (128-byte rows, the same TMA boxes / UMMA layout as every other wgrad).
#include <cuda_fp16.h>
#include <float.h>
#include <stdlib.h>

#include "launch.h"
#include "ptx.cuh"

namespace b2n {

constexpr int kStemC = 32;   // channels of the weight gradient's reduction per tap (12 real)
constexpr int kStemCStored = 12;  // channels of the fp32 copy actually stored: the weight gradient's
                                  // TMA map declares 12 channels and requests 32 -- the other 20 are
                                  // zero-filled by the TMA unit instead of being written and read
constexpr int kStemC16 = 16; // channels of the FP16 (forward) pair

// x: NCHW (N,3,H,W), H and W even, fp32 or uint8 pixels.  Outputs NHWC: the (hi, lo) FP16 pair
// (N,H/2,W/2,16) for the forward conv and, when xs32 is given, the TF32-rounded fp32 copy
// (N,H/2,W/2,32) the weight gradient reads.  uint8 pixels are exact in fp16: their lo plane is
// identically zero and is neither written nor (lo_nonzero stays 0) read by the conv.
template <typename T>
__device__ __forceinline__ float2 load_px2(const T* p);
template <>
__device__ __forceinline__ float2 load_px2<float>(const float* p) {
  return *reinterpret_cast<const float2*>(p);
}
template <>
__device__ __forceinline__ float2 load_px2<unsigned char>(const unsigned char* p) {
  const uchar2 v = *reinterpret_cast<const uchar2*>(p);
  return make_float2(static_cast<float>(v.x), static_cast<float>(v.y));
}

template <typename T>
__global__ void stem_pack_input_kernel(const T* __restrict__ x, uint4* __restrict__ xs_h,
                                       uint4* __restrict__ xs_l, float4* __restrict__ xs32,
                                       int* __restrict__ lo_nonzero, int N, int H, int W) {
  constexpr bool kExact = sizeof(T) == 1;
  const int H2 = H >> 1, W2 = W >> 1;
  const size_t total = static_cast<size_t>(N) * H2 * W2;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += stride) {
    const int j = static_cast<int>(t % W2);
    const int i = static_cast<int>((t / W2) % H2);
    const int n = static_cast<int>(t / (static_cast<size_t>(W2) * H2));
    float v[16];
#pragma unroll
    for (int k = 12; k < 16; ++k) v[k] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const T* plane = x + (static_cast<size_t>(n) * 3 + c) * H * W;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const float2 p = load_px2<T>(plane + static_cast<size_t>(2 * i + dy) * W + 2 * j);
        v[(dy * 2 + 0) * 3 + c] = p.x;
        v[(dy * 2 + 1) * 3 + c] = p.y;
      }
    }
    uint4 ph[2], pl[2];
    __half2* h2 = reinterpret_cast<__half2*>(ph);
    __half2* l2 = reinterpret_cast<__half2*>(pl);
#pragma unroll
    for (int k = 0; k < 8; ++k) split_f16(v[2 * k], v[2 * k + 1], h2[k], l2[k]);
    uint4* dh = xs_h + t * (kStemC16 / 8);
    dh[0] = ph[0]; dh[1] = ph[1];
    if (!kExact) {
      // integer-valued images (uint8 patches cast to float, dataset.py:65-67) are exact in fp16:
      // tell the stem conv it may skip the all-zero lo plane
      if (lo_nonzero != nullptr && ((pl[0].x | pl[0].y | pl[0].z | pl[0].w | pl[1].x | pl[1].y |
                                     pl[1].z | pl[1].w) & 0x7fff7fffu) != 0)
        atomicOr(lo_nonzero, 1);
      uint4* dl = xs_l + t * (kStemC16 / 8);
      dl[0] = pl[0]; dl[1] = pl[1];
    }
    if (xs32 != nullptr) {
      float4* d = xs32 + t * (kStemCStored / 4);
      d[0] = make_float4(tf32_rn(v[0]), tf32_rn(v[1]), tf32_rn(v[2]), tf32_rn(v[3]));
      d[1] = make_float4(tf32_rn(v[4]), tf32_rn(v[5]), tf32_rn(v[6]), tf32_rn(v[7]));
      d[2] = make_float4(tf32_rn(v[8]), tf32_rn(v[9]), tf32_rn(v[10]), tf32_rn(v[11]));
    }
  }
}

// w: (K,3,7,7) -> (hi, lo) FP16 pair ws: [K][16 taps][16].  unpack = the transpose map for grads
// (of the 32-channel fp32 layout the weight gradient runs in).
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, __half* __restrict__ ws_h,
                                        __half* __restrict__ ws_l, int K) {
  const int total = K * 16 * kStemC16;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int ch = t % kStemC16;
    const int tap = (t / kStemC16) % 16;
    const int k = t / (kStemC16 * 16);
    float v = 0.f;
    if (ch < 12) {
      const int c = ch % 3, dd = ch / 3, dy = dd >> 1, dx = dd & 1;
      const int r = 2 * (tap >> 2) + dy - 1, s = 2 * (tap & 3) + dx - 1;
      if (r >= 0 && r < 7 && s >= 0 && s < 7) v = w[((k * 3 + c) * 7 + r) * 7 + s];
    }
    const __half h = __float2half_rn(v);
    ws_h[t] = h;
    ws_l[t] = __float2half_rn(v - __half2float(h));
  }
}
__global__ void stem_unpack_wgrad_kernel(const float* __restrict__ dws, float* __restrict__ dw,
                                         int K, int accumulate, int planes) {
  const int total = K * 3 * 49;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int s = t % 7, r = (t / 7) % 7, c = (t / 49) % 3, k = t / 147;
    const int tr = (r + 1) >> 1, dy = (r + 1) & 1, ts = (s + 1) >> 1, dx = (s + 1) & 1;
    const int src = (k * 16 + tr * 4 + ts) * kStemC + (dy * 2 + dx) * 3 + c;
    float v = dws[src];
    for (int pl = 1; pl < planes; ++pl) v += dws[static_cast<size_t>(pl) * K * 16 * kStemC + src];
    dw[t] = accumulate ? dw[t] + v : v;
  }
}

template <typename T>
static int launch_stem_pack_input_t(const T* x, __half* xs_h, __half* xs_l, float* xs32,
                                    int* lo_nonzero, int N, int H, int W, cudaStream_t stream) {
  if ((H | W) & 1) return set_error("stem_pack_input: H and W must be even (got %dx%d)", H, W);
  if (sizeof(T) != 1 && xs_l == nullptr) return set_error("stem_pack_input: xs_l is null");
  const size_t total = static_cast<size_t>(N) * (H / 2) * (W / 2);
  size_t blocks = (total + 127) / 128;
  const size_t cap = static_cast<size_t>(device_sm_count()) * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  stem_pack_input_kernel<T><<<(unsigned)blocks, 128, 0, stream>>>(
      x, reinterpret_cast<uint4*>(xs_h), reinterpret_cast<uint4*>(xs_l),
      reinterpret_cast<float4*>(xs32), lo_nonzero, N, H, W);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("stem_pack_input: %s", cudaGetErrorString(e));
  return 0;
}
int launch_stem_pack_input(const float* x, __half* xs_h, __half* xs_l, float* xs32,
                           int* lo_nonzero, int N, int H, int W, cudaStream_t stream) {
  return launch_stem_pack_input_t<float>(x, xs_h, xs_l, xs32, lo_nonzero, N, H, W, stream);
}
int launch_stem_pack_input_u8(const unsigned char* x, __half* xs_h, float* xs32, int N, int H, int W,
                              cudaStream_t stream) {
  return launch_stem_pack_input_t<unsigned char>(x, xs_h, nullptr, xs32, nullptr, N, H, W, stream);
}
int launch_stem_pack_weight(const float* w, __half* ws_h, __half* ws_l, int K,
                            cudaStream_t stream) {
  stem_pack_weight_kernel<<<64, 256, 0, stream>>>(w, ws_h, ws_l, K);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("stem_pack_weight: %s", cudaGetErrorString(e));
  return 0;
}
int launch_stem_unpack_wgrad(const float* dws, float* dw, int K, int accumulate, int planes,
                             cudaStream_t stream) {
  if (planes < 1) return set_error("stem_unpack_wgrad: planes must be >= 1");
  stem_unpack_wgrad_kernel<<<37, 256, 0, stream>>>(dws, dw, K, accumulate, planes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("stem_unpack_wgrad: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------- BN + ReLU + maxpool 3x3/s2/p1
// a[n,p,q,c] = max_{3x3 window} relu(scale*y + shift); idx = r*3+s of the first maximum
// (same tie rule as ATen's max_pool2d, torchvision resnet.py:271).  idx may be null (eval).
__global__ void bn_relu_maxpool_kernel(const float4* __restrict__ y, const float* __restrict__ scale,
                                       const float* __restrict__ shift, float4* __restrict__ a32,
                                       uint2* __restrict__ a_h, uint2* __restrict__ a_l,
                                       uchar4* __restrict__ idx, int N, int H, int W, int P, int Q,
                                       int C4) {
  const size_t total = static_cast<size_t>(N) * P * Q * C4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += stride) {
    const int c4 = static_cast<int>(t % C4);
    size_t u = t / C4;
    const int q = static_cast<int>(u % Q); u /= Q;
    const int p = static_cast<int>(u % P);
    const int n = static_cast<int>(u / P);
    const float4 sc = *reinterpret_cast<const float4*>(scale + 4 * c4);
    const float4 sh = *reinterpret_cast<const float4*>(shift + 4 * c4);
    float best[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    unsigned char bi[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * p - 1 + r;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = 2 * q - 1 + s;
        if (w < 0 || w >= W) continue;
        const float4 v = y[((static_cast<size_t>(n) * H + h) * W + w) * C4 + c4];
        const float z[4] = {fmaxf(fmaf(v.x, sc.x, sh.x), 0.f), fmaxf(fmaf(v.y, sc.y, sh.y), 0.f),
                            fmaxf(fmaf(v.z, sc.z, sh.z), 0.f), fmaxf(fmaf(v.w, sc.w, sh.w), 0.f)};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (z[k] > best[k]) { best[k] = z[k]; bi[k] = static_cast<unsigned char>(r * 3 + s); }
      }
    }
    if (a_h != nullptr) {
      uint2 ph, pl;
      __half2* h2 = reinterpret_cast<__half2*>(&ph);
      __half2* l2 = reinterpret_cast<__half2*>(&pl);
      split_f16(best[0], best[1], h2[0], l2[0]);
      split_f16(best[2], best[3], h2[1], l2[1]);
      a_h[t] = ph;
      a_l[t] = pl;
    }
    if (a32 != nullptr)
      a32[t] = make_float4(tf32_rn(best[0]), tf32_rn(best[1]), tf32_rn(best[2]), tf32_rn(best[3]));
    if (idx != nullptr) idx[t] = make_uchar4(bi[0], bi[1], bi[2], bi[3]);
  }
}

// Same computation as a bulk-copy pipeline: a persistent block walks pooled rows (n, p); one
// thread streams the (up to) three rows of y the row's windows cover into a double-buffered
// shared-memory slot with a single 1-D bulk copy (TMA), so the next row's ~86 KB are in flight
// while this one is reduced.  Adjacent pooled rows share one y row: it is re-read from L2.
constexpr int kPoolRowThreads = 512;

__global__ void __launch_bounds__(kPoolRowThreads, 1)
bn_relu_maxpool_rows_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                            const float* __restrict__ shift, float4* __restrict__ a32,
                            uint2* __restrict__ a_h, uint2* __restrict__ a_l,
                            uchar4* __restrict__ idx, int N, int H, int W, int P, int Q, int C) {
  extern __shared__ uint8_t pool_smem_raw[];
  uint8_t* smem = pool_smem_raw + ((128u - (smem_u32(pool_smem_raw) & 127u)) & 127u);
  const int tid = threadIdx.x;
  const int C4 = C >> 2;
  const uint32_t row_bytes = static_cast<uint32_t>(W) * C * 4;
  const uint32_t slot_bytes = 3 * row_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 2 * slot_bytes);
  const int nbands = N * P;
  if (tid == 0) {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int band, int slot) {
    const int n = band / P, p = band - n * P;
    const int h_lo = 2 * p - 1 < 0 ? 0 : 2 * p - 1;
    const int h_hi = 2 * p + 1 > H - 1 ? H - 1 : 2 * p + 1;
    const uint32_t bytes = static_cast<uint32_t>(h_hi - h_lo + 1) * row_bytes;
    mbar_arrive_expect_tx(&full_bar[slot], bytes);
    bulk_load_1d(smem + slot * slot_bytes, y + (static_cast<size_t>(n) * H + h_lo) * W * C, bytes,
                 &full_bar[slot]);
  };
  int k = 0;
  if (tid == 0 && static_cast<int>(blockIdx.x) < nbands) issue(blockIdx.x, 0);
  for (int band = blockIdx.x; band < nbands; band += gridDim.x, ++k) {
    const int slot = k & 1;
    if (tid == 0 && band + static_cast<int>(gridDim.x) < nbands) issue(band + gridDim.x, slot ^ 1);
    mbar_wait(&full_bar[slot], (k >> 1) & 1);
    const int n = band / P, p = band - n * P;
    const int h_lo = 2 * p - 1 < 0 ? 0 : 2 * p - 1;
    const float4* ys = reinterpret_cast<const float4*>(smem + slot * slot_bytes);
    for (int i = tid; i < Q * C4; i += kPoolRowThreads) {
      const int q = i / C4, c4 = i - q * C4;
      const float4 sc = *reinterpret_cast<const float4*>(scale + 4 * c4);
      const float4 sh = *reinterpret_cast<const float4*>(shift + 4 * c4);
      float best[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
      unsigned char bi[4] = {0, 0, 0, 0};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int h = 2 * p - 1 + r;
        if (h < 0 || h >= H) continue;
#pragma unroll
        for (int sx = 0; sx < 3; ++sx) {
          const int w = 2 * q - 1 + sx;
          if (w < 0 || w >= W) continue;
          const float4 v = ys[((h - h_lo) * W + w) * C4 + c4];
          const float z[4] = {fmaxf(fmaf(v.x, sc.x, sh.x), 0.f), fmaxf(fmaf(v.y, sc.y, sh.y), 0.f),
                              fmaxf(fmaf(v.z, sc.z, sh.z), 0.f), fmaxf(fmaf(v.w, sc.w, sh.w), 0.f)};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (z[kk] > best[kk]) { best[kk] = z[kk]; bi[kk] = static_cast<unsigned char>(r * 3 + sx); }
        }
      }
      const size_t t = (static_cast<size_t>(band) * Q) * C4 + i;
      if (a_h != nullptr) {
        uint2 ph, pl;
        __half2* h2 = reinterpret_cast<__half2*>(&ph);
        __half2* l2 = reinterpret_cast<__half2*>(&pl);
        split_f16(best[0], best[1], h2[0], l2[0]);
        split_f16(best[2], best[3], h2[1], l2[1]);
        a_h[t] = ph;
        a_l[t] = pl;
      }
      if (a32 != nullptr)
        a32[t] = make_float4(tf32_rn(best[0]), tf32_rn(best[1]), tf32_rn(best[2]), tf32_rn(best[3]));
      if (idx != nullptr) idx[t] = make_uchar4(bi[0], bi[1], bi[2], bi[3]);
    }
    __syncthreads();  // everyone is done with this slot before it is refilled
  }
}

// Gradient of maxpool + ReLU w.r.t. the BN output: gz[n,h,w,c] = [scale*y+shift > 0] *
// sum over the (<= 4) windows whose recorded argmax is (h,w) of ga.
__global__ void maxpool_relu_bwd_kernel(const float4* __restrict__ ga, const uchar4* __restrict__ idx,
                                        const float4* __restrict__ y, const float* __restrict__ scale,
                                        const float* __restrict__ shift, float4* __restrict__ gz,
                                        int N, int H, int W, int P, int Q, int C4) {
  const size_t total = static_cast<size_t>(N) * H * W * C4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += stride) {
    const int c4 = static_cast<int>(t % C4);
    size_t u = t / C4;
    const int w = static_cast<int>(u % W); u /= W;
    const int h = static_cast<int>(u % H);
    const int n = static_cast<int>(u / H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int p_lo = h >> 1, p_hi = (h + 1) >> 1;  // windows with 2p-1 <= h <= 2p+1
    const int q_lo = w >> 1, q_hi = (w + 1) >> 1;
    for (int p = p_lo; p <= p_hi; ++p) {
      if (p >= P) continue;
      const int r = h - (2 * p - 1);
      for (int q = q_lo; q <= q_hi; ++q) {
        if (q >= Q) continue;
        const int s = w - (2 * q - 1);
        const unsigned char code = static_cast<unsigned char>(r * 3 + s);
        const size_t o = ((static_cast<size_t>(n) * P + p) * Q + q) * C4 + c4;
        const uchar4 id = idx[o];
        const float4 g = ga[o];
        if (id.x == code) acc[0] += g.x;
        if (id.y == code) acc[1] += g.y;
        if (id.z == code) acc[2] += g.z;
        if (id.w == code) acc[3] += g.w;
      }
    }
    const float4 v = y[t];
    const float4 sc = *reinterpret_cast<const float4*>(scale + 4 * c4);
    const float4 sh = *reinterpret_cast<const float4*>(shift + 4 * c4);
    float4 o;
    o.x = fmaf(v.x, sc.x, sh.x) > 0.f ? acc[0] : 0.f;
    o.y = fmaf(v.y, sc.y, sh.y) > 0.f ? acc[1] : 0.f;
    o.z = fmaf(v.z, sc.z, sh.z) > 0.f ? acc[2] : 0.f;
    o.w = fmaf(v.w, sc.w, sh.w) > 0.f ? acc[3] : 0.f;
    gz[t] = o;
  }
}

static unsigned grid_for(size_t total, int threads) {
  size_t blocks = (total + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(device_sm_count()) * elementwise_blocks_per_sm();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

int launch_bn_relu_maxpool(const float* y, const float* scale, const float* shift, float* a32,
                           __half* a_h, __half* a_l, unsigned char* idx, int N, int H, int W,
                           int C, cudaStream_t stream) {
  if (C % 4 != 0) return set_error("bn_relu_maxpool: C %% 4 != 0");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const size_t total = static_cast<size_t>(N) * P * Q * (C / 4);
  // pipelined variant when two slots of three y rows fit in shared memory
  const size_t smem = 2 * 3 * static_cast<size_t>(W) * C * 4 + 16 + 128;
  if ((static_cast<size_t>(W) * C * 4) % 16 == 0 && smem <= 227 * 1024 && getenv("B2N_NO_POOL_PIPE") == nullptr) {
    static PerDeviceMax configured;
    int dev;
    if (smem > 48 * 1024 && configured.needs((int)smem, &dev)) {
      cudaError_t e = cudaFuncSetAttribute(bn_relu_maxpool_rows_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return set_error("bn_relu_maxpool: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      configured.set(dev, (int)smem);
    }
    int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
    if (per_sm > 2) per_sm = 2;
    int grid = device_sm_count() * per_sm;
    if (grid > N * P) grid = N * P;
    bn_relu_maxpool_rows_kernel<<<grid, kPoolRowThreads, smem, stream>>>(
        y, scale, shift, reinterpret_cast<float4*>(a32), reinterpret_cast<uint2*>(a_h),
        reinterpret_cast<uint2*>(a_l), reinterpret_cast<uchar4*>(idx), N, H, W, P, Q, C);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("bn_relu_maxpool: %s", cudaGetErrorString(e));
    return 0;
  }
  bn_relu_maxpool_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(y), scale, shift, reinterpret_cast<float4*>(a32),
      reinterpret_cast<uint2*>(a_h), reinterpret_cast<uint2*>(a_l), reinterpret_cast<uchar4*>(idx),
      N, H, W, P, Q, C / 4);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("bn_relu_maxpool: %s", cudaGetErrorString(e));
  return 0;
}

int launch_maxpool_relu_bwd(const float* ga, const unsigned char* idx, const float* y,
                            const float* scale, const float* shift, float* gz, int N, int H, int W,
                            int C, cudaStream_t stream) {
  if (C % 4 != 0) return set_error("maxpool_relu_bwd: C %% 4 != 0");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const size_t total = static_cast<size_t>(N) * H * W * (C / 4);
  maxpool_relu_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(ga), reinterpret_cast<const uchar4*>(idx),
      reinterpret_cast<const float4*>(y), scale, shift, reinterpret_cast<float4*>(gz), N, H, W, P, Q,
      C / 4);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("maxpool_relu_bwd: %s", cudaGetErrorString(e));
  return 0;
}

// --------------------------------- fused maxpool + ReLU + BatchNorm backward (the stem's tail)
// The stem's BN output is 1/3 of all conv-output elements of the trunk, so materialising
// gz = d(maxpool o relu) and then running the generic two-pass BN backward over it costs three
// extra sweeps of that tensor.  Here gz never exists in memory: a persistent block walks bands of
// two image rows (2b, 2b+1); for every band one thread issues three 1-D bulk copies (TMA) into a
// double-buffered shared-memory slot -- the band of y, and the pooled gradients and recorded
// argmax codes of pooled rows b and b+1 (the only rows whose 3x3/s2 windows reach the band) --
// so the next band's ~100 KB are in flight while this one is swept.  gz is rebuilt per position
// from its (at most four) candidate windows in maxpool_relu_bwd_kernel's order:
//   reduce: sums[0][c] += sum gz', sums[1][c] += sum gz' * xhat       (gz' = gz * [scale*y+shift > 0])
//   apply : dy = gamma*invstd*(gz' - sums0/rows - xhat*sums1/rows); dgamma = sums1, dbeta = sums0
// The reduce pass issues one atomic per channel per block.
constexpr int kBandThreads = 512;

template <bool APPLY, bool ROUND>
__global__ void __launch_bounds__(kBandThreads, 1)
pool_bn_bwd_band_kernel(const float4* __restrict__ ga, const uchar4* __restrict__ idx,
                        const float4* __restrict__ y, const float* __restrict__ scale,
                        const float* __restrict__ shift, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                        double* __restrict__ sums, float4* __restrict__ dy, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, int N, int H, int W, int P, int Q, int C,
                        double inv_count, int accumulate) {
  extern __shared__ uint8_t band_smem_raw[];
  uint8_t* smem = band_smem_raw + ((128u - (smem_u32(band_smem_raw) & 127u)) & 127u);
  const int tid = threadIdx.x;
  const int C4 = C >> 2;
  const int c4 = tid % C4;  // constant per thread: kBandThreads % C4 == 0
  const int HB = (H + 1) >> 1;
  const int nbands = N * HB;
  // slot layout: y band [2][W][C] fp32 | pooled gradients [2][Q][C] fp32 | argmax codes [2][Q][C] u8
  const uint32_t y_bytes = 2u * W * C * 4, g_bytes = 2u * Q * C * 4, i_bytes = 2u * Q * C;
  const uint32_t slot_bytes = y_bytes + g_bytes + i_bytes;  // multiple of 16 (C % 4 == 0, checked)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 2 * slot_bytes);

  const float4 sc = *reinterpret_cast<const float4*>(scale + 4 * c4);
  const float4 sh = *reinterpret_cast<const float4*>(shift + 4 * c4);
  const float4 mu = *reinterpret_cast<const float4*>(mean + 4 * c4);
  const float4 is = *reinterpret_cast<const float4*>(invstd + 4 * c4);
  float4 m1 = make_float4(0, 0, 0, 0), m2 = m1, gi = m1;
  if (APPLY) {
    const int c = 4 * c4;
    m1 = make_float4(static_cast<float>(sums[c] * inv_count), static_cast<float>(sums[c + 1] * inv_count),
                     static_cast<float>(sums[c + 2] * inv_count), static_cast<float>(sums[c + 3] * inv_count));
    m2 = make_float4(static_cast<float>(sums[C + c] * inv_count), static_cast<float>(sums[C + c + 1] * inv_count),
                     static_cast<float>(sums[C + c + 2] * inv_count), static_cast<float>(sums[C + c + 3] * inv_count));
    const float4 ga4 = *reinterpret_cast<const float4*>(gamma + c);
    gi = make_float4(ga4.x * is.x, ga4.y * is.y, ga4.z * is.z, ga4.w * is.w);
    if (blockIdx.x == 0 && dgamma != nullptr) {
      for (int k = tid; k < C; k += kBandThreads) {
        dbeta[k] = (accumulate ? dbeta[k] : 0.f) + static_cast<float>(sums[k]);
        dgamma[k] = (accumulate ? dgamma[k] : 0.f) + static_cast<float>(sums[C + k]);
      }
    }
  }
  float4 s1 = make_float4(0, 0, 0, 0), s2 = s1;

  if (tid == 0) {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();

  // one thread streams a band into a slot; rows / pooled rows beyond the image are not copied
  auto issue = [&](int band, int slot) {
    const int n = band / HB, hb = band - n * HB, h0 = 2 * hb;
    const uint32_t rows = h0 + 1 < H ? 2u : 1u, prow = hb + 1 < P ? 2u : 1u;
    uint8_t* dst = smem + slot * slot_bytes;
    const size_t po = (static_cast<size_t>(n) * P + hb) * Q * C;
    mbar_arrive_expect_tx(&full_bar[slot], rows * (y_bytes / 2) + prow * (g_bytes / 2) + prow * (i_bytes / 2));
    bulk_load_1d(dst, reinterpret_cast<const float*>(y) + (static_cast<size_t>(n) * H + h0) * W * C,
                 rows * (y_bytes / 2), &full_bar[slot]);
    bulk_load_1d(dst + y_bytes, reinterpret_cast<const float*>(ga) + po, prow * (g_bytes / 2), &full_bar[slot]);
    bulk_load_1d(dst + y_bytes + g_bytes, reinterpret_cast<const unsigned char*>(idx) + po,
                 prow * (i_bytes / 2), &full_bar[slot]);
  };

  int k = 0;
  if (tid == 0 && static_cast<int>(blockIdx.x) < nbands) issue(blockIdx.x, 0);
  for (int band = blockIdx.x; band < nbands; band += gridDim.x, ++k) {
    const int slot = k & 1;
    // slot ^ 1 was released by the __syncthreads that ended the previous iteration
    if (tid == 0 && band + static_cast<int>(gridDim.x) < nbands) issue(band + gridDim.x, slot ^ 1);
    mbar_wait(&full_bar[slot], (k >> 1) & 1);
    const int n = band / HB, hb = band - n * HB, h0 = 2 * hb;
    const float4* y_s = reinterpret_cast<const float4*>(smem + slot * slot_bytes);
    const float4* g_s = reinterpret_cast<const float4*>(smem + slot * slot_bytes + y_bytes);
    const uchar4* id_s = reinterpret_cast<const uchar4*>(smem + slot * slot_bytes + y_bytes + g_bytes);
    const int band_n = (h0 + 1 < H ? 2 : 1) * W * C4;  // float4 elements of this band
    const size_t t0 = (static_cast<size_t>(n) * H + h0) * W * C4;
#pragma unroll 2
    for (int i = tid; i < band_n; i += kBandThreads) {
      const int hh = i / (W * C4);
      const int w = (i - hh * W * C4) / C4;
      const float4 v = y_s[i];
      // branch-free gather over the four candidate windows (pr, q) in maxpool_relu_bwd_kernel's
      // order; candidates that do not exist get a code no recorded argmax can equal
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const int q_lo = w >> 1, q_hi = (w + 1) >> 1;
#pragma unroll
      for (int cand = 0; cand < 4; ++cand) {
        const int pr = cand >> 1;            // row h0: window row hb only; row h0+1: hb, hb+1
        const int q = (cand & 1) ? q_hi : q_lo;
        const bool valid = pr <= hh && hb + pr < P && (!(cand & 1) || q_hi != q_lo) && q < Q;
        const int code = valid ? (hh + 1 - 2 * pr) * 3 + (w - (2 * q - 1)) : 254;
        const int sidx = ((valid ? pr : 0) * Q + (q < Q ? q : Q - 1)) * C4 + c4;
        const uchar4 id = id_s[sidx];
        const float4 gq = g_s[sidx];
        acc[0] += id.x == code ? gq.x : 0.f;
        acc[1] += id.y == code ? gq.y : 0.f;
        acc[2] += id.z == code ? gq.z : 0.f;
        acc[3] += id.w == code ? gq.w : 0.f;
      }
      float4 g = make_float4(acc[0], acc[1], acc[2], acc[3]);
      g.x = fmaf(v.x, sc.x, sh.x) > 0.f ? g.x : 0.f;
      g.y = fmaf(v.y, sc.y, sh.y) > 0.f ? g.y : 0.f;
      g.z = fmaf(v.z, sc.z, sh.z) > 0.f ? g.z : 0.f;
      g.w = fmaf(v.w, sc.w, sh.w) > 0.f ? g.w : 0.f;
      const float4 xh = make_float4((v.x - mu.x) * is.x, (v.y - mu.y) * is.y, (v.z - mu.z) * is.z,
                                    (v.w - mu.w) * is.w);
      if (APPLY) {
        float4 o = make_float4(gi.x * (g.x - m1.x - xh.x * m2.x), gi.y * (g.y - m1.y - xh.y * m2.y),
                               gi.z * (g.z - m1.z - xh.z * m2.z), gi.w * (g.w - m1.w - xh.w * m2.w));
        if (ROUND) o = make_float4(tf32_rn(o.x), tf32_rn(o.y), tf32_rn(o.z), tf32_rn(o.w));
        dy[t0 + i] = o;
      } else {
        s1.x += g.x; s1.y += g.y; s1.z += g.z; s1.w += g.w;
        s2.x += g.x * xh.x; s2.y += g.y * xh.y; s2.z += g.z * xh.z; s2.w += g.w * xh.w;
      }
    }
    __syncthreads();  // everyone is done with this slot before it is refilled
  }
  if (!APPLY) {
    // block reduction over the kBandThreads / C4 threads that share a channel group
    float* red = reinterpret_cast<float*>(smem);  // [kBandThreads][8]
    float* dst = red + tid * 8;
    dst[0] = s1.x; dst[1] = s1.y; dst[2] = s1.z; dst[3] = s1.w;
    dst[4] = s2.x; dst[5] = s2.y; dst[6] = s2.z; dst[7] = s2.w;
    __syncthreads();
    for (int j = tid; j < 2 * C; j += kBandThreads) {
      const int which = j / C, c = j - which * C;
      float acc = 0.f;
      for (int t = c >> 2; t < kBandThreads; t += C4) acc += red[t * 8 + which * 4 + (c & 3)];
      atomicAdd(&sums[which * C + c], static_cast<double>(acc));
    }
  }
}

template <bool APPLY, bool ROUND>
static int launch_band(const float* ga, const unsigned char* idx, const float* y, const float* scale,
                       const float* shift, const float* mean, const float* invstd, const float* gamma,
                       double* sums, float* dy, float* dgamma, float* dbeta, int N, int H, int W,
                       int C, int accumulate, cudaStream_t stream) {
  const int C4 = C / 4;
  if (C % 16 != 0 || kBandThreads % C4 != 0)
    return set_error("pool_bn_bwd: unsupported C=%d", C);
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  // bulk copies need 16-byte sizes: one image row of y, one pooled row of gradients / codes
  if ((W * C * 4) % 16 != 0 || (Q * C) % 16 != 0)
    return set_error("pool_bn_bwd: row sizes must be multiples of 16 bytes (W=%d, C=%d)", W, C);
  const size_t slot = static_cast<size_t>(2) * W * C * 4 + static_cast<size_t>(2) * Q * C * 5;
  size_t smem = 2 * slot + 16 + 128;
  if (smem < kBandThreads * 8 * sizeof(float) + 128) smem = kBandThreads * 8 * sizeof(float) + 128;
  if (smem > 227 * 1024) return set_error("pool_bn_bwd: image row too wide (W=%d, C=%d)", W, C);
  auto kern = pool_bn_bwd_band_kernel<APPLY, ROUND>;
  static PerDeviceMax configured;
  int dev;
  if (smem > 48 * 1024 && configured.needs((int)smem, &dev)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("pool_bn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured.set(dev, (int)smem);
  }
  const int nbands = N * ((H + 1) / 2);
  int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = device_sm_count() * per_sm;
  if (grid > nbands) grid = nbands;
  const double inv_count = 1.0 / (static_cast<double>(N) * H * W);
  kern<<<grid, kBandThreads, smem, stream>>>(
      reinterpret_cast<const float4*>(ga), reinterpret_cast<const uchar4*>(idx),
      reinterpret_cast<const float4*>(y), scale, shift, mean, invstd, gamma, sums,
      reinterpret_cast<float4*>(dy), dgamma, dbeta, N, H, W, P, Q, C, inv_count, accumulate);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("pool_bn_bwd: %s", cudaGetErrorString(e));
  return 0;
}

int launch_pool_bn_bwd_reduce(const float* ga, const unsigned char* idx, const float* y,
                              const float* scale, const float* shift, const float* mean,
                              const float* invstd, double* sums, int N, int H, int W, int C,
                              cudaStream_t stream) {
  return launch_band<false, false>(ga, idx, y, scale, shift, mean, invstd, nullptr, sums, nullptr,
                                   nullptr, nullptr, N, H, W, C, 0, stream);
}

int launch_pool_bn_bwd_apply(const float* ga, const unsigned char* idx, const float* y,
                             const float* scale, const float* shift, const float* mean,
                             const float* invstd, const float* gamma, const double* sums, float* dy,
                             float* dgamma, float* dbeta, int N, int H, int W, int C, int round_tf32,
                             int accumulate, cudaStream_t stream) {
  double* s = const_cast<double*>(sums);  // only read when APPLY
  if (round_tf32)
    return launch_band<true, true>(ga, idx, y, scale, shift, mean, invstd, gamma, s, dy, dgamma, dbeta,
                                   N, H, W, C, accumulate, stream);
  return launch_band<true, false>(ga, idx, y, scale, shift, mean, invstd, gamma, s, dy, dgamma, dbeta,
                                  N, H, W, C, accumulate, stream);
}

// ------------------------------------------------------------ global avg-pool
// a: [N][HW][C] -> e: [N][C]  (AdaptiveAvgPool2d(1) + flatten, torchvision resnet.py:278-279)
__global__ void avgpool_fwd_kernel(const __half* __restrict__ a_h, const __half* __restrict__ a_l,
                                   float* __restrict__ e, int HW, int C) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < HW; ++i) {
      const size_t o = (static_cast<size_t>(n) * HW + i) * C + c;
      acc += __half2float(a_h[o]) + __half2float(a_l[o]);
    }
    e[static_cast<size_t>(n) * C + c] = acc / static_cast<float>(HW);
  }
}
// gate (optional): the pooled activation itself -- its ReLU gate is applied here, at the
// producer of the gradient, so no consumer has to read a mask tensor
__global__ void avgpool_bwd_kernel(const float* __restrict__ ge, const float* __restrict__ gate,
                                   float* __restrict__ g, int HW, int C) {
  const int n = blockIdx.x;
  const float inv = 1.f / static_cast<float>(HW);
  for (int t = threadIdx.x; t < HW * C; t += blockDim.x) {
    const size_t i = static_cast<size_t>(n) * HW * C + t;
    const float v = ge[static_cast<size_t>(n) * C + (t % C)] * inv;
    g[i] = gate == nullptr || gate[i] > 0.f ? v : 0.f;
  }
}
int launch_avgpool_fwd(const __half* a_h, const __half* a_l, float* e, int N, int HW, int C,
                       cudaStream_t stream) {
  avgpool_fwd_kernel<<<N, 256, 0, stream>>>(a_h, a_l, e, HW, C);
  cudaError_t er = cudaGetLastError();
  if (er != cudaSuccess) return set_error("avgpool_fwd: %s", cudaGetErrorString(er));
  return 0;
}
int launch_avgpool_bwd(const float* ge, const float* gate, float* g, int N, int HW, int C,
                       cudaStream_t stream) {
  avgpool_bwd_kernel<<<N, 256, 0, stream>>>(ge, gate, g, HW, C);
  cudaError_t er = cudaGetLastError();
  if (er != cudaSuccess) return set_error("avgpool_bwd: %s", cudaGetErrorString(er));
  return 0;
}

}  // namespace b2n
