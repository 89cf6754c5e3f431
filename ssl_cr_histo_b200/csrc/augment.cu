// Weak / strong augmentation of uint8 patch batches on the GPU.
//
// The reference builds its consistency-training views on the CPU, one PIL / numpy image at a time:
// TransformFix (dataset.py:663-677) = RandomHorizontalFlip + RandomCrop for the weak view and the
// same followed by RandAugment(n, m=10) (models/randaugment.py:112-144) for the strong view, whose
// nine operations call albumentations 0.1.8 / imgaug 0.4 / scikit-image 0.15 / OpenCV
// (requirements.txt:10,128,369,241) -- including per-pixel Python loops (models/randaugment.py:36-39,
// dataset.py:93-96).  Here every operation is one launch over a whole (N,3,H,W) uint8 batch with
// per-image parameters (device arrays), so the loader only has to ship raw uint8 patches.
//
// These are HBM-bound byte kernels: one thread handles one pixel position of one image (all three
// channel planes), consecutive threads consecutive pixels of a row, so every plane access is
// coalesced.  All arithmetic that decides a rounding is written with explicit round-to-nearest
// single operations (__fmul_rn / __fadd_rn: no FMA contraction) or in integers / doubles so that the
// CPU restatement (oracle/ref_augment.py) is reproduced bit for bit.
//
// `apply` (optional, per image): 0 = copy the image through unchanged (RandAugment draws a different
// operation list for every image; albumentations transforms fire with probability 0.5).
#include <math.h>

#include "launch.h"

namespace b2n {

namespace {

__device__ __forceinline__ int reflect101(int p, int n) {
  // OpenCV BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

__device__ __forceinline__ unsigned char sat_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

struct Geo {
  int N, H, W;  // output (and, unless stated otherwise, input) geometry; 3 channel planes
};

__device__ __forceinline__ bool pixel_of(const Geo& g, int& n, int& y, int& x) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(g.N) * g.H * g.W) return false;
  x = static_cast<int>(t % g.W);
  const long long u = t / g.W;
  y = static_cast<int>(u % g.H);
  n = static_cast<int>(u / g.H);
  return true;
}

}  // namespace

// ---------------------------------------------------------------- flip + crop (weak view)
// transforms.RandomHorizontalFlip then transforms.RandomCrop (dataset.py:668-669): the crop window
// (top, left) is taken from the flipped image.
__global__ void aug_flip_crop_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
                                     const int* __restrict__ top, const int* __restrict__ left,
                                     const int* __restrict__ flip, Geo g, int Hs, int Ws) {
  int n, y, x;
  if (!pixel_of(g, n, y, x)) return;
  const int sy = top[n] + y;
  const int sx = flip[n] ? Ws - 1 - (left[n] + x) : left[n] + x;
  const size_t sp = static_cast<size_t>(Hs) * Ws, dp = static_cast<size_t>(g.H) * g.W;
  const unsigned char* s = src + static_cast<size_t>(n) * 3 * sp + static_cast<size_t>(sy) * Ws + sx;
  unsigned char* d = dst + static_cast<size_t>(n) * 3 * dp + static_cast<size_t>(y) * g.W + x;
  d[0] = s[0];
  d[dp] = s[sp];
  d[2 * dp] = s[2 * sp];
}

// ---------------------------------------------------------------- brightness / contrast
// albumentations brightness_contrast_adjust: img.astype(float32) * alpha + offset, clipped to
// [0, 255], cast (truncated) to uint8; offset = beta * mean(img) or beta * 255, formed by the caller.
__global__ void aug_brightness_contrast_kernel(const unsigned char* __restrict__ src,
                                               unsigned char* __restrict__ dst,
                                               const float* __restrict__ alpha,
                                               const float* __restrict__ offset,
                                               const int* __restrict__ apply, Geo g) {
  int n, y, x;
  if (!pixel_of(g, n, y, x)) return;
  const size_t plane = static_cast<size_t>(g.H) * g.W;
  const size_t o = static_cast<size_t>(n) * 3 * plane + static_cast<size_t>(y) * g.W + x;
  const bool on = apply == nullptr || apply[n] != 0;
  const float a = alpha[n], b = offset[n];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const unsigned char v = src[o + c * plane];
    float f = __fadd_rn(__fmul_rn(static_cast<float>(v), a), b);
    f = fminf(fmaxf(f, 0.f), 255.f);
    dst[o + c * plane] = on ? static_cast<unsigned char>(f) : v;
  }
}

// per-image mean over all three planes (the `beta * np.mean(img)` reference of older albumentations)
__global__ void aug_image_mean_kernel(const unsigned char* __restrict__ src, float* __restrict__ mean,
                                      int count) {
  __shared__ unsigned long long part[8];
  const unsigned char* s = src + static_cast<size_t>(blockIdx.x) * count;
  unsigned long long acc = 0;
  for (int i = threadIdx.x; i < count; i += blockDim.x) acc += s[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += part[w];
    mean[blockIdx.x] = static_cast<float>(static_cast<double>(t) / count);   // exact integer sum
  }
}

// ---------------------------------------------------------------- hue / saturation / value shift
// albumentations 0.1.8 shift_hsv on uint8: cv2 RGB2HSV (8-bit: H in [0,180), fixed-point tables),
// hue += dh with one wrap (h < 0: +180, h > 180: -180), saturation / value += ds / dv clipped to
// [0,255], cv2 HSV2RGB (8-bit: float32 sector formula, truncated).  dh/ds/dv are the integer shifts
// cv2.add applies (the caller rounds the sampled floats).
__global__ void aug_hsv_shift_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
                                     const int* __restrict__ dh, const int* __restrict__ ds,
                                     const int* __restrict__ dv, const int* __restrict__ apply, Geo g) {
  int n, y, x;
  if (!pixel_of(g, n, y, x)) return;
  const size_t plane = static_cast<size_t>(g.H) * g.W;
  const size_t o = static_cast<size_t>(n) * 3 * plane + static_cast<size_t>(y) * g.W + x;
  const int r = src[o], gr = src[o + plane], b = src[o + 2 * plane];
  if (apply != nullptr && apply[n] == 0) {
    dst[o] = r; dst[o + plane] = gr; dst[o + 2 * plane] = b;
    return;
  }
  // RGB -> HSV, OpenCV's 8-bit integer path (hsv_shift = 12, rounded division tables)
  const int v = max(max(r, gr), b), vmin = min(min(r, gr), b);
  const int diff = v - vmin;
  const int sdiv = v > 0 ? static_cast<int>(rint((255 << 12) / static_cast<double>(v))) : 0;
  const int hdiv = diff > 0 ? static_cast<int>(rint((180 << 12) / (6.0 * diff))) : 0;
  int s = (diff * sdiv + (1 << 11)) >> 12;
  int h;
  if (v == r) h = gr - b;
  else if (v == gr) h = b - r + 2 * diff;
  else h = r - gr + 4 * diff;
  h = (h * hdiv + (1 << 11)) >> 12;
  if (h < 0) h += 180;
  // the shifts
  h += dh[n];
  if (h < 0) h += 180;
  if (h > 180) h -= 180;
  h = min(max(h, 0), 255);                 // .astype(uint8) of an in-range value
  s = min(max(s + ds[n], 0), 255);
  const int vv = min(max(v + dv[n], 0), 255);
  // HSV -> RGB, OpenCV's 8-bit path: float32, no FMA, truncation
  const float fs = __fmul_rn(static_cast<float>(s), 1.f / 255.f);
  const float fv = __fmul_rn(static_cast<float>(vv), 1.f / 255.f);
  float fr, fg, fb;
  if (s == 0) {
    fr = fg = fb = fv;
  } else {
    float fh = __fmul_rn(static_cast<float>(h), 6.f / 180.f);
    int sector = static_cast<int>(floorf(fh));
    fh = __fsub_rn(fh, static_cast<float>(sector));
    if (static_cast<unsigned>(sector) >= 6u) { sector = 0; fh = 0.f; }
    float tab[4];
    tab[0] = fv;
    tab[1] = __fmul_rn(fv, __fsub_rn(1.f, fs));
    tab[2] = __fmul_rn(fv, __fsub_rn(1.f, __fmul_rn(fs, fh)));
    tab[3] = __fmul_rn(fv, __fsub_rn(1.f, __fmul_rn(fs, __fsub_rn(1.f, fh))));
    // sector_data[sector] = indices of (b, g, r) into tab
    const int code = sector == 0 ? 0x130 : sector == 1 ? 0x102 : sector == 2 ? 0x301
                   : sector == 3 ? 0x021 : sector == 4 ? 0x013 : 0x210;
    fb = tab[(code >> 8) & 3];
    fg = tab[(code >> 4) & 3];
    fr = tab[code & 3];
  }
  dst[o] = sat_u8(static_cast<int>(floorf(__fmul_rn(fr, 255.f))));
  dst[o + plane] = sat_u8(static_cast<int>(floorf(__fmul_rn(fg, 255.f))));
  dst[o + 2 * plane] = sat_u8(static_cast<int>(floorf(__fmul_rn(fb, 255.f))));
}

// ---------------------------------------------------------------- additive Gaussian noise
// imgaug AdditiveGaussianNoise(per_channel=False): one noise value per pixel shared by the channels;
// out = clip(round(img + noise)).  The noise field (N,1,H,W) fp32 is an input (any generator).
__global__ void aug_add_noise_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
                                     const float* __restrict__ noise, const int* __restrict__ apply,
                                     Geo g) {
  int n, y, x;
  if (!pixel_of(g, n, y, x)) return;
  const size_t plane = static_cast<size_t>(g.H) * g.W;
  const size_t o = static_cast<size_t>(n) * 3 * plane + static_cast<size_t>(y) * g.W + x;
  const bool on = apply == nullptr || apply[n] != 0;
  const float z = noise[static_cast<size_t>(n) * plane + static_cast<size_t>(y) * g.W + x];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const unsigned char v = src[o + c * plane];
    const float f = rintf(__fadd_rn(static_cast<float>(v), z));       // round half to even (np.round)
    dst[o + c * plane] = on ? sat_u8(static_cast<int>(fminf(fmaxf(f, -1.f), 256.f))) : v;
  }
}

// ---------------------------------------------------------------- box blur
// albumentations Blur -> cv2.blur(img, (k, k)): normalised box filter, BORDER_REFLECT_101, integer
// window sum scaled by 1/k^2 and rounded to nearest.  k odd, 1 <= k <= 7 (k = 1: copy).
__global__ void aug_box_blur_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
                                    const int* __restrict__ ksize, const int* __restrict__ apply, Geo g) {
  int n, y, x;
  if (!pixel_of(g, n, y, x)) return;
  const size_t plane = static_cast<size_t>(g.H) * g.W;
  const size_t base = static_cast<size_t>(n) * 3 * plane;
  const size_t o = base + static_cast<size_t>(y) * g.W + x;
  const int k = (apply == nullptr || apply[n] != 0) ? ksize[n] : 1;
  if (k <= 1) {
    dst[o] = src[o]; dst[o + plane] = src[o + plane]; dst[o + 2 * plane] = src[o + 2 * plane];
    return;
  }
  const int r = k >> 1;
  int acc[3] = {0, 0, 0};
  for (int dy = -r; dy <= r; ++dy) {
    const int yy = reflect101(y + dy, g.H);
    for (int dx = -r; dx <= r; ++dx) {
      const size_t p = base + static_cast<size_t>(yy) * g.W + reflect101(x + dx, g.W);
      acc[0] += src[p]; acc[1] += src[p + plane]; acc[2] += src[p + 2 * plane];
    }
  }
  const double scale = 1.0 / (k * k);
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[o + c * plane] = sat_u8(static_cast<int>(rint(acc[c] * scale)));
}

// ---------------------------------------------------------------- H&E-DAB stain jitter
// colour_augmentation (models/randaugment.py:17-48, dataset.py:75-106): rgb2hed, add one offset per
// stain channel to every pixel (the reference does it with a Python loop over all pixels), hed2rgb,
// (x * 255).astype(uint8).  scikit-image 0.15 formulas: stains = -log(rgb/255 + 2) . inv(M);
// rgb' = clip(exp(-stains' . M) - 2, -1, 1).  Double precision, like the reference; the final cast
// truncates and wraps negatives modulo 256 as numpy's float64 -> uint8 conversion does.
__global__ void aug_hed_jitter_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
                                      const float* __restrict__ delta /* [N][3] */,
                                      const int* __restrict__ apply, Geo g) {
  int n, y, x;
  if (!pixel_of(g, n, y, x)) return;
  const size_t plane = static_cast<size_t>(g.H) * g.W;
  const size_t o = static_cast<size_t>(n) * 3 * plane + static_cast<size_t>(y) * g.W + x;
  if (apply != nullptr && apply[n] == 0) {
    dst[o] = src[o]; dst[o + plane] = src[o + plane]; dst[o + 2 * plane] = src[o + 2 * plane];
    return;
  }
  // rgb_from_hed (Ruifrok & Johnston) and its inverse, rows = stains
  const double M[3][3] = {{0.65, 0.70, 0.29}, {0.07, 0.99, 0.11}, {0.27, 0.57, 0.78}};
  // (scipy.linalg.inv(rgb_from_hed), as skimage.color computes hed_from_rgb)
  const double I[3][3] = {{1.8779827368521356, -1.0076786862855642, -0.5561158181996246},
                          {-0.06590806222356334, 1.1347303724996625, -0.13552179862837116},
                          {-0.6019073634392891, -0.4804141884970579, 1.5735880719641926}};
  double l[3], st[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) l[c] = -log(src[o + c * plane] / 255.0 + 2.0);
#pragma unroll
  for (int j = 0; j < 3; ++j)
    st[j] = l[0] * I[0][j] + l[1] * I[1][j] + l[2] * I[2][j] + static_cast<double>(delta[3 * n + j]);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double lg = -(st[0] * M[0][c] + st[1] * M[1][c] + st[2] * M[2][c]);
    double z = exp(lg) - 2.0;
    z = fmin(fmax(z, -1.0), 1.0);
    const long long q = static_cast<long long>(z * 255.0);      // truncation toward zero
    dst[o + c * plane] = static_cast<unsigned char>(q & 0xFF);  // numpy wraps negatives
  }
}

// ---------------------------------------------------------------- affine warp, bicubic
// cv2.warpAffine / cv2.resize with INTER_CUBIC (interpolation=2 in the reference's Rotate,
// ShiftScaleRotate, RandomScale, Resize) and BORDER_REFLECT_101.  minv [N][6] maps an output pixel
// centre (x, y) to source coordinates sx = m0 x + m1 y + m2, sy = m3 x + m4 y + m5.  Cubic kernel
// with a = -0.75 (OpenCV's); coefficients in fp32 on the exact coordinate (OpenCV quantises it to
// 1/32 pixel and uses fixed-point weights, so it differs from this by a few grey levels at most).
// Source and destination sizes may differ.  clamp_border != 0 uses edge replication (cv2.resize).
__device__ __forceinline__ void cubic_weights(float t, float w[4]) {
  const float A = -0.75f;
  w[0] = ((A * (t + 1.f) - 5.f * A) * (t + 1.f) + 8.f * A) * (t + 1.f) - 4.f * A;
  w[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  w[2] = ((A + 2.f) * (1.f - t) - (A + 3.f)) * (1.f - t) * (1.f - t) + 1.f;
  w[3] = 1.f - w[0] - w[1] - w[2];
}

__global__ void aug_warp_affine_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
                                       const float* __restrict__ minv, const int* __restrict__ apply,
                                       Geo g, int Hs, int Ws, int clamp_border) {
  int n, y, x;
  if (!pixel_of(g, n, y, x)) return;
  const size_t sp = static_cast<size_t>(Hs) * Ws, dp = static_cast<size_t>(g.H) * g.W;
  const unsigned char* s = src + static_cast<size_t>(n) * 3 * sp;
  unsigned char* d = dst + static_cast<size_t>(n) * 3 * dp + static_cast<size_t>(y) * g.W + x;
  if (apply != nullptr && apply[n] == 0) {   // only meaningful when the sizes agree
    const size_t p = static_cast<size_t>(min(y, Hs - 1)) * Ws + min(x, Ws - 1);
    d[0] = s[p]; d[dp] = s[p + sp]; d[2 * dp] = s[p + 2 * sp];
    return;
  }
  const float* m = minv + 6 * n;
  const float sx = fmaf(m[0], static_cast<float>(x), fmaf(m[1], static_cast<float>(y), m[2]));
  const float sy = fmaf(m[3], static_cast<float>(x), fmaf(m[4], static_cast<float>(y), m[5]));
  const int ix = static_cast<int>(floorf(sx)), iy = static_cast<int>(floorf(sy));
  float wx[4], wy[4];
  cubic_weights(sx - static_cast<float>(ix), wx);
  cubic_weights(sy - static_cast<float>(iy), wy);
  int xs[4], ys[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int px = ix - 1 + k, py = iy - 1 + k;
    xs[k] = clamp_border ? min(max(px, 0), Ws - 1) : reflect101(px, Ws);
    ys[k] = clamp_border ? min(max(py, 0), Hs - 1) : reflect101(py, Hs);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float row = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) row += wx[k] * static_cast<float>(s[c * sp + static_cast<size_t>(ys[j]) * Ws + xs[k]]);
      acc += wy[j] * row;
    }
    d[c * dp] = sat_u8(static_cast<int>(rintf(fminf(fmaxf(acc, -1.f), 256.f))));
  }
}

// ---------------------------------------------------------------- launchers
static int geo_grid(const Geo& g, unsigned* blocks) {
  const long long total = static_cast<long long>(g.N) * g.H * g.W;
  if (g.N < 0 || g.H <= 0 || g.W <= 0) return 1;
  *blocks = static_cast<unsigned>((total + 255) / 256);
  return 0;
}
#define B2N_AUG_CHECK(name)                                                      \
  {                                                                              \
    cudaError_t e = cudaGetLastError();                                          \
    if (e != cudaSuccess) return set_error(name ": %s", cudaGetErrorString(e)); \
  }

int launch_aug_flip_crop(const unsigned char* src, unsigned char* dst, const int* top, const int* left,
                         const int* flip, int N, int Hs, int Ws, int H, int W, cudaStream_t stream) {
  Geo g{N, H, W};
  unsigned blocks;
  if (geo_grid(g, &blocks) || H > Hs || W > Ws) return set_error("aug_flip_crop: bad geometry");
  if (N == 0) return 0;
  aug_flip_crop_kernel<<<blocks, 256, 0, stream>>>(src, dst, top, left, flip, g, Hs, Ws);
  B2N_AUG_CHECK("aug_flip_crop");
  return 0;
}
int launch_aug_brightness_contrast(const unsigned char* src, unsigned char* dst, const float* alpha,
                                   const float* offset, const int* apply, int N, int H, int W,
                                   cudaStream_t stream) {
  Geo g{N, H, W};
  unsigned blocks;
  if (geo_grid(g, &blocks)) return set_error("aug_brightness_contrast: bad geometry");
  if (N == 0) return 0;
  aug_brightness_contrast_kernel<<<blocks, 256, 0, stream>>>(src, dst, alpha, offset, apply, g);
  B2N_AUG_CHECK("aug_brightness_contrast");
  return 0;
}
int launch_aug_image_mean(const unsigned char* src, float* mean, int N, int H, int W, cudaStream_t stream) {
  if (N < 0 || H <= 0 || W <= 0) return set_error("aug_image_mean: bad geometry");
  if (N == 0) return 0;
  aug_image_mean_kernel<<<N, 256, 0, stream>>>(src, mean, 3 * H * W);
  B2N_AUG_CHECK("aug_image_mean");
  return 0;
}
int launch_aug_hsv_shift(const unsigned char* src, unsigned char* dst, const int* dh, const int* ds,
                         const int* dv, const int* apply, int N, int H, int W, cudaStream_t stream) {
  Geo g{N, H, W};
  unsigned blocks;
  if (geo_grid(g, &blocks)) return set_error("aug_hsv_shift: bad geometry");
  if (N == 0) return 0;
  aug_hsv_shift_kernel<<<blocks, 256, 0, stream>>>(src, dst, dh, ds, dv, apply, g);
  B2N_AUG_CHECK("aug_hsv_shift");
  return 0;
}
int launch_aug_add_noise(const unsigned char* src, unsigned char* dst, const float* noise, const int* apply,
                         int N, int H, int W, cudaStream_t stream) {
  Geo g{N, H, W};
  unsigned blocks;
  if (geo_grid(g, &blocks)) return set_error("aug_add_noise: bad geometry");
  if (N == 0) return 0;
  aug_add_noise_kernel<<<blocks, 256, 0, stream>>>(src, dst, noise, apply, g);
  B2N_AUG_CHECK("aug_add_noise");
  return 0;
}
int launch_aug_box_blur(const unsigned char* src, unsigned char* dst, const int* ksize, const int* apply,
                        int N, int H, int W, cudaStream_t stream) {
  Geo g{N, H, W};
  unsigned blocks;
  if (geo_grid(g, &blocks)) return set_error("aug_box_blur: bad geometry");
  if (N == 0) return 0;
  aug_box_blur_kernel<<<blocks, 256, 0, stream>>>(src, dst, ksize, apply, g);
  B2N_AUG_CHECK("aug_box_blur");
  return 0;
}
int launch_aug_hed_jitter(const unsigned char* src, unsigned char* dst, const float* delta, const int* apply,
                          int N, int H, int W, cudaStream_t stream) {
  Geo g{N, H, W};
  unsigned blocks;
  if (geo_grid(g, &blocks)) return set_error("aug_hed_jitter: bad geometry");
  if (N == 0) return 0;
  aug_hed_jitter_kernel<<<blocks, 256, 0, stream>>>(src, dst, delta, apply, g);
  B2N_AUG_CHECK("aug_hed_jitter");
  return 0;
}
int launch_aug_warp_affine(const unsigned char* src, unsigned char* dst, const float* minv, const int* apply,
                           int N, int Hs, int Ws, int H, int W, int clamp_border, cudaStream_t stream) {
  Geo g{N, H, W};
  unsigned blocks;
  if (geo_grid(g, &blocks) || Hs <= 0 || Ws <= 0) return set_error("aug_warp_affine: bad geometry");
  if (apply != nullptr && (Hs != H || Ws != W))
    return set_error("aug_warp_affine: a per-image apply mask needs equal source and output sizes");
  if (N == 0) return 0;
  aug_warp_affine_kernel<<<blocks, 256, 0, stream>>>(src, dst, minv, apply, g, Hs, Ws, clamp_border);
  B2N_AUG_CHECK("aug_warp_affine");
  return 0;
}

}  // namespace b2n
