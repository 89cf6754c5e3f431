"""CPU oracle (numpy) for the weak / strong augmentation operations.  TEST INFRASTRUCTURE ONLY.

Restates, operation by operation, what the reference's view pipeline computes per image:
TransformFix (dataset.py:663-677) and the RandAugment pool (models/randaugment.py:51-126).  The
arithmetic itself lives in third-party dependencies that are absent from /root/reference and from
this image -- albumentations=0.1.8, imgaug=0.4.0, scikit-image=0.15.0 (requirements.txt:10,128,369)
-- and in OpenCV (requirements.txt:241; cv2 4.13 is installed here).  Their published algorithms
are restated below; tests/test_augment.py pins every function that bottoms out in OpenCV against
the installed cv2 (blur, RGB<->HSV, warpAffine / resize) and the rest against closed forms.  Parity
status of the albumentations / imgaug / skimage formulas: unpinned (no copy to run).

All functions take and return uint8 arrays of shape (N, 3, H, W) with per-image parameters.
"""
from __future__ import annotations

import numpy as np

# ------------------------------------------------------------------ flip + crop


def flip_crop(x, top, left, flip, H, W):
    """transforms.RandomHorizontalFlip + transforms.RandomCrop (dataset.py:668-669)."""
    out = np.empty((x.shape[0], 3, H, W), np.uint8)
    for n in range(x.shape[0]):
        img = x[n, :, :, ::-1] if flip[n] else x[n]
        out[n] = img[:, top[n]:top[n] + H, left[n]:left[n] + W]
    return out


# ------------------------------------------------------------------ brightness / contrast


def brightness_contrast(x, alpha, offset, apply=None):
    """albumentations brightness_contrast_adjust (RandomBrightnessContrast,
    models/randaugment.py:91-101): float32 img * alpha + offset, clip, truncate."""
    out = x.copy()
    for n in range(x.shape[0]):
        if apply is not None and not apply[n]:
            continue
        f = x[n].astype(np.float32) * np.float32(alpha[n]) + np.float32(offset[n])
        out[n] = np.clip(f, 0, 255).astype(np.uint8)
    return out


def image_mean(x):
    return x.reshape(x.shape[0], -1).astype(np.float64).mean(1).astype(np.float32)


# ------------------------------------------------------------------ HSV shift
_SDIV = np.zeros(256, np.int64)
_HDIV = np.zeros(256, np.int64)
for _i in range(1, 256):
    _SDIV[_i] = int(np.rint((255 << 12) / (1.0 * _i)))
    _HDIV[_i] = int(np.rint((180 << 12) / (6.0 * _i)))


def rgb2hsv_u8(rgb):
    """OpenCV cvtColor(COLOR_RGB2HSV) on uint8 (imgproc/src/color_hsv: RGB2HSV_b, hsv_shift 12).
    rgb: (..., 3) uint8 -> (..., 3) uint8 with H in [0, 180)."""
    c = rgb.astype(np.int64)
    r, g, b = c[..., 0], c[..., 1], c[..., 2]
    v = np.maximum(np.maximum(r, g), b)
    diff = v - np.minimum(np.minimum(r, g), b)
    s = (diff * _SDIV[v] + (1 << 11)) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * _HDIV[diff] + (1 << 11)) >> 12
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def hsv2rgb_u8(hsv):
    """OpenCV cvtColor(COLOR_HSV2RGB) on uint8: HSV2RGB_native's sector formula in float32, result
    truncated.  Against the installed cv2 4.13: identical on 99.995 % of uniformly random HSV
    triples; on HSV triples that came from integer RGB (where the exact result is an integer) cv2
    is one level higher on ~1.4 % of the values -- never more than one level (tests/test_augment.py).
    The reference's own OpenCV (3.4.2 / 4.2, requirements.txt:241-242) cannot be run here."""
    f = np.float32
    h = hsv[..., 0].astype(f) * f(6.0 / 180.0)
    s = hsv[..., 1].astype(f) * f(1.0 / 255.0)
    v = hsv[..., 2].astype(f) * f(1.0 / 255.0)
    sector = np.floor(h).astype(np.int64)
    hf = (h - sector.astype(f)).astype(f)
    over = sector >= 6
    sector = np.where(over, 0, sector)
    hf = np.where(over, f(0), hf)
    one = f(1)
    tabs = np.stack([v, v * (one - s), v * (one - s * hf), v * (one - s * (one - hf))], -1)
    sd = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])
    idx = sd[sector]
    b = np.take_along_axis(tabs, idx[..., 0:1], -1)[..., 0]
    g = np.take_along_axis(tabs, idx[..., 1:2], -1)[..., 0]
    r = np.take_along_axis(tabs, idx[..., 2:3], -1)[..., 0]
    grey = hsv[..., 1] == 0
    r, g, b = (np.where(grey, v, t) for t in (r, g, b))
    out = np.stack([r, g, b], -1) * f(255)
    return np.clip(np.floor(out), 0, 255).astype(np.uint8)


def hsv_shift(x, dh, ds, dv, apply=None):
    """albumentations 0.1.8 shift_hsv (HueSaturationValue, models/randaugment.py:51-57)."""
    out = x.copy()
    for n in range(x.shape[0]):
        if apply is not None and not apply[n]:
            continue
        hsv = rgb2hsv_u8(np.transpose(x[n], (1, 2, 0))).astype(np.int64)
        h = hsv[..., 0] + int(dh[n])
        h = np.where(h < 0, h + 180, h)
        h = np.where(h > 180, h - 180, h)
        s = np.clip(hsv[..., 1] + int(ds[n]), 0, 255)
        v = np.clip(hsv[..., 2] + int(dv[n]), 0, 255)
        rgb = hsv2rgb_u8(np.stack([np.clip(h, 0, 255), s, v], -1).astype(np.uint8))
        out[n] = np.transpose(rgb, (2, 0, 1))
    return out


# ------------------------------------------------------------------ noise / blur


def add_noise(x, noise, apply=None):
    """imgaug AdditiveGaussianNoise(per_channel=False) (IAAAdditiveGaussianNoise,
    models/randaugment.py:59-63): noise (N,1,H,W) float32 shared by the channels."""
    out = x.copy()
    for n in range(x.shape[0]):
        if apply is not None and not apply[n]:
            continue
        f = np.rint(x[n].astype(np.float32) + noise[n].astype(np.float32))
        out[n] = np.clip(f, 0, 255).astype(np.uint8)
    return out


def _reflect101(p, n):
    if n == 1:
        return np.zeros_like(p)
    p = np.asarray(p).copy()
    while ((p < 0) | (p >= n)).any():
        p = np.where(p < 0, -p, p)
        p = np.where(p >= n, 2 * n - 2 - p, p)
    return p


def box_blur(x, ksize, apply=None):
    """albumentations Blur -> cv2.blur(img, (k, k)) (models/randaugment.py:85-89)."""
    out = x.copy()
    N, _, H, W = x.shape
    for n in range(N):
        k = int(ksize[n]) if (apply is None or apply[n]) else 1
        if k <= 1:
            continue
        r = k // 2
        ys = _reflect101(np.arange(-r, H + r), H)
        xs = _reflect101(np.arange(-r, W + r), W)
        pad = x[n][:, ys][:, :, xs].astype(np.int64)
        acc = np.zeros((3, H, W), np.int64)
        for dy in range(k):
            for dx in range(k):
                acc += pad[:, dy:dy + H, dx:dx + W]
        out[n] = np.clip(np.rint(acc * (1.0 / (k * k))), 0, 255).astype(np.uint8)
    return out


# ------------------------------------------------------------------ H&E-DAB jitter
RGB_FROM_HED = np.array([[0.65, 0.70, 0.29], [0.07, 0.99, 0.11], [0.27, 0.57, 0.78]])
# scipy.linalg.inv(RGB_FROM_HED) -- how skimage.color builds hed_from_rgb
HED_FROM_RGB = np.array([[1.8779827368521356, -1.0076786862855642, -0.5561158181996246],
                         [-0.06590806222356334, 1.1347303724996625, -0.13552179862837116],
                         [-0.6019073634392891, -0.4804141884970579, 1.5735880719641926]])


def hed_jitter(x, delta, apply=None):
    """colour_augmentation (models/randaugment.py:17-48) with scikit-image 0.15's separate_stains /
    combine_stains: stains = -log(rgb + 2) . hed_from_rgb; + per-channel offsets; rgb' =
    rescale_intensity(exp(-stains . rgb_from_hed) - 2, in_range=(-1, 1)) = clip to [-1, 1];
    (rgb' * 255).astype('uint8') (truncation; negatives wrap modulo 256 like numpy on x86)."""
    out = x.copy()
    for n in range(x.shape[0]):
        if apply is not None and not apply[n]:
            continue
        rgb = np.transpose(x[n], (1, 2, 0)).astype(np.float64) / 255.0 + 2.0
        stains = (-np.log(rgb)).reshape(-1, 3) @ HED_FROM_RGB
        stains = stains + np.asarray(delta[n], np.float32).astype(np.float64)[None, :]
        z = np.exp(-(stains @ RGB_FROM_HED)) - 2.0
        z = np.clip(z, -1.0, 1.0).reshape(x.shape[2], x.shape[3], 3)
        q = np.trunc(z * 255.0).astype(np.int64) & 0xFF
        out[n] = np.transpose(q.astype(np.uint8), (2, 0, 1))
    return out


# ------------------------------------------------------------------ bicubic affine warp


def _cubic_weights(t):
    A = np.float32(-0.75)
    t = t.astype(np.float32)
    one = np.float32(1)
    w0 = ((A * (t + one) - np.float32(5) * A) * (t + one) + np.float32(8) * A) * (t + one) - np.float32(4) * A
    w1 = ((A + np.float32(2)) * t - (A + np.float32(3))) * t * t + one
    u = one - t
    w2 = ((A + np.float32(2)) * u - (A + np.float32(3))) * u * u + one
    w3 = one - w0 - w1 - w2
    return np.stack([w0, w1, w2, w3], -1)


def warp_affine(x, minv, H, W, apply=None, clamp_border=False):
    """cv2.warpAffine(..., flags=INTER_CUBIC | WARP_INVERSE_MAP, borderMode=BORDER_REFLECT_101) with
    exact (unquantised) coordinates; minv (N, 6) maps output (x, y) to source (sx, sy).
    clamp_border: edge replication instead (cv2.resize)."""
    N, _, Hs, Ws = x.shape
    out = np.empty((N, 3, H, W), np.uint8)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    for n in range(N):
        if apply is not None and not apply[n]:
            out[n] = x[n][:, :H, :W]
            continue
        m = np.asarray(minv[n], np.float32)
        # fmaf chains, as the kernel: m0*x + (m1*y + m2) with single roundings
        sx = (m[0].astype(np.float64) * xx + (m[1].astype(np.float64) * yy + m[2]).astype(np.float32)).astype(np.float32)
        sy = (m[3].astype(np.float64) * xx + (m[4].astype(np.float64) * yy + m[5]).astype(np.float32)).astype(np.float32)
        ix, iy = np.floor(sx).astype(np.int64), np.floor(sy).astype(np.int64)
        wx, wy = _cubic_weights(sx - ix.astype(np.float32)), _cubic_weights(sy - iy.astype(np.float32))
        acc = np.zeros((3, H, W), np.float64)
        for j in range(4):
            py = iy - 1 + j
            py = np.clip(py, 0, Hs - 1) if clamp_border else _reflect101(py, Hs)
            row = np.zeros((3, H, W), np.float64)
            for k in range(4):
                px = ix - 1 + k
                px = np.clip(px, 0, Ws - 1) if clamp_border else _reflect101(px, Ws)
                row += wx[..., k].astype(np.float64) * x[n][:, py, px]
            acc += wy[..., j].astype(np.float64) * row
        out[n] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return out


def rotation_matrix_inv(cx, cy, angle_deg, scale, dx=0.0, dy=0.0):
    """Inverse of cv2.getRotationMatrix2D((cx, cy), angle, scale) followed by a (dx, dy) shift -- the
    matrix albumentations' rotate / shift_scale_rotate hand to cv2.warpAffine -- as the six
    output->source coefficients warp_affine takes."""
    a = np.deg2rad(angle_deg)
    al, be = scale * np.cos(a), scale * np.sin(a)
    M = np.array([[al, be, (1 - al) * cx - be * cy + dx], [-be, al, be * cx + (1 - al) * cy + dy], [0, 0, 1.0]])
    return np.linalg.inv(M)[:2].reshape(6).astype(np.float32)


def resize_matrix_inv(src_h, src_w, dst_h, dst_w):
    """cv2.resize's pixel-centre mapping sx = (x + 0.5) * src_w / dst_w - 0.5."""
    fx, fy = src_w / dst_w, src_h / dst_h
    return np.array([fx, 0, 0.5 * fx - 0.5, 0, fy, 0.5 * fy - 0.5], np.float32)
