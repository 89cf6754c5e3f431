"""Weak / strong views for consistency training, built on the GPU (csrc/augment.cu).

The reference's loaders make both views per image on the CPU (dataset.py:663-677):

    weak   = RandomHorizontalFlip + RandomCrop(image_size)
    strong = RandomHorizontalFlip + RandomCrop(image_size) + RandAugment(n=N, m=10)

with RandAugment (models/randaugment.py:112-144) drawing ``n`` operations per image from a pool of
nine (HSV, Noise, Scale_Resize_Crop, Shift_Scale_Rotate, Color, Blur, Brightness, Contrast,
Rotate_Crop) at magnitude ``val = v/30 * (max - min) + min``, ``v = randint(1, m)``.  Here the host
only *draws the parameters* (same distributions, per image) and every operation runs as one kernel
launch over the whole uint8 (N,3,H,W) batch; the result feeds the trunk's uint8 input path
directly.  The low-level functions take explicit per-image parameters, which is what the parity
tests drive against the CPU restatement of the same operations.

There is no CPU fallback: tensors must live on a CUDA (sm_100a) device.
"""
from __future__ import annotations

import math
import random as _random
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import call


def _check(x: torch.Tensor) -> torch.Tensor:
    _lib.require_device(x, "image batch")
    if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[1] != 3:
        raise RuntimeError("expected a uint8 (N,3,H,W) batch, got %s %s" % (x.dtype, tuple(x.shape)))
    return x.contiguous()


def _dev(values, dtype, x):
    if isinstance(values, torch.Tensor):
        return values.to(device=x.device, dtype=dtype).contiguous()
    return torch.as_tensor(np.asarray(values), dtype=dtype).to(x.device)


def _mask(apply, x):
    return None if apply is None else _dev(apply, torch.int32, x)


# ---------------------------------------------------------------------------- operations
def flip_crop(x, top, left, flip, size: Tuple[int, int]) -> torch.Tensor:
    """RandomHorizontalFlip + RandomCrop (dataset.py:668-669): crop window from the flipped image."""
    x = _check(x)
    N, _, Hs, Ws = x.shape
    H, W = size
    out = torch.empty(N, 3, H, W, device=x.device, dtype=torch.uint8)
    call("b2n_aug_flip_crop", x, out, _dev(top, torch.int32, x), _dev(left, torch.int32, x),
         _dev(flip, torch.int32, x), N, Hs, Ws, H, W)
    return out


def brightness_contrast(x, alpha, beta, apply=None, beta_by_max: bool = False) -> torch.Tensor:
    """albumentations RandomBrightnessContrast arithmetic (models/randaugment.py:91-101):
    uint8(clip(float32(x) * alpha + beta * ref)), ref = the image mean (albumentations 0.1.x) or 255
    (``beta_by_max``, later versions)."""
    x = _check(x)
    N, _, H, W = x.shape
    beta = _dev(beta, torch.float32, x)
    if beta_by_max:
        offset = beta * 255.0
    else:
        mean = torch.empty(N, device=x.device, dtype=torch.float32)
        call("b2n_aug_image_mean", x, mean, N, H, W)
        offset = beta * mean
    out = torch.empty_like(x)
    call("b2n_aug_brightness_contrast", x, out, _dev(alpha, torch.float32, x), offset.contiguous(),
         _mask(apply, x), N, H, W)
    return out


def hsv_shift(x, dh, ds, dv, apply=None) -> torch.Tensor:
    """albumentations 0.1.8 shift_hsv (models/randaugment.py:51-57) with integer shifts."""
    x = _check(x)
    N, _, H, W = x.shape
    out = torch.empty_like(x)
    call("b2n_aug_hsv_shift", x, out, _dev(dh, torch.int32, x), _dev(ds, torch.int32, x),
         _dev(dv, torch.int32, x), _mask(apply, x), N, H, W)
    return out


def add_noise(x, noise, apply=None) -> torch.Tensor:
    """imgaug AdditiveGaussianNoise (models/randaugment.py:59-63); noise: (N,1,H,W) float32."""
    x = _check(x)
    N, _, H, W = x.shape
    noise = _dev(noise, torch.float32, x)
    if noise.numel() != N * H * W:
        raise RuntimeError("noise must hold one value per pixel: (N,1,H,W)")
    out = torch.empty_like(x)
    call("b2n_aug_add_noise", x, out, noise, _mask(apply, x), N, H, W)
    return out


def box_blur(x, ksize, apply=None) -> torch.Tensor:
    """albumentations Blur = cv2.blur (models/randaugment.py:85-89); ksize odd, <= 7."""
    x = _check(x)
    N, _, H, W = x.shape
    ks = np.asarray(ksize.cpu() if isinstance(ksize, torch.Tensor) else ksize)
    if ((ks % 2 == 0) | (ks < 1) | (ks > 7)).any():
        raise RuntimeError("box_blur: kernel sizes must be odd and in [1, 7]")
    out = torch.empty_like(x)
    call("b2n_aug_box_blur", x, out, _dev(ks, torch.int32, x), _mask(apply, x), N, H, W)
    return out


def hed_jitter(x, delta, apply=None) -> torch.Tensor:
    """colour_augmentation (models/randaugment.py:17-48): delta (N,3) stain offsets."""
    x = _check(x)
    N, _, H, W = x.shape
    out = torch.empty_like(x)
    call("b2n_aug_hed_jitter", x, out, _dev(delta, torch.float32, x).view(-1), _mask(apply, x), N, H, W)
    return out


def warp_affine(x, minv, size: Optional[Tuple[int, int]] = None, apply=None,
                clamp_border: bool = False) -> torch.Tensor:
    """Bicubic cv2.warpAffine / cv2.resize; minv (N,6) maps output (x, y) to source coordinates."""
    x = _check(x)
    N, _, Hs, Ws = x.shape
    H, W = size if size is not None else (Hs, Ws)
    out = torch.empty(N, 3, H, W, device=x.device, dtype=torch.uint8)
    call("b2n_aug_warp_affine", x, out, _dev(minv, torch.float32, x).view(-1), _mask(apply, x), N, Hs, Ws,
         H, W, 1 if clamp_border else 0)
    return out


# ---------------------------------------------------------------------------- matrices
def rotation_matrix_inv(cx, cy, angle_deg, scale, dx=0.0, dy=0.0, flip_h=False, flip_v=False,
                        width=0, height=0):
    """Output->source coefficients of cv2.getRotationMatrix2D((cx,cy), angle, scale) + (dx,dy) shift
    (albumentations rotate / shift_scale_rotate), optionally preceded by a flip of the source."""
    a = math.radians(angle_deg)
    al, be = scale * math.cos(a), scale * math.sin(a)
    M = np.array([[al, be, (1 - al) * cx - be * cy + dx], [-be, al, be * cx + (1 - al) * cy + dy], [0, 0, 1.0]])
    F = np.eye(3)
    if flip_h:
        F = np.array([[-1.0, 0, width - 1], [0, 1, 0], [0, 0, 1]]) @ F
    if flip_v:
        F = np.array([[1.0, 0, 0], [0, -1, height - 1], [0, 0, 1]]) @ F
    # destination = M . (flipped source); source = F^-1 . M^-1 . destination (a flip is its own inverse)
    return (F @ np.linalg.inv(M))[:2].reshape(6).astype(np.float32)


def resize_matrix_inv(src_h, src_w, dst_h, dst_w):
    fx, fy = src_w / dst_w, src_h / dst_h
    return np.array([fx, 0, 0.5 * fx - 0.5, 0, fy, 0.5 * fy - 0.5], np.float32)


# ---------------------------------------------------------------------------- RandAugment
_POOL = (("HSV", -1, 1), ("Noise", 0, 0.15), ("Scale_Resize_Crop", 0.8, 1.2),
         ("Shift_Scale_Rotate", 0.01, 0.1), ("Color", -0.035, 0.035), ("Blur_img", 0, 2),
         ("Brightness", -0.2, 0.2), ("Contrast", -0.2, 0.2), ("Rotate_Crop", -90, 90))   # :112-123


class RandAugment:
    """models/randaugment.py:126-144 over a batch: per image ``n`` operations drawn with replacement
    from the pool, magnitude from ``randint(1, m)``; the albumentations transforms inside each
    operation fire with their default probability 0.5 and draw their own parameters -- all of it
    sampled on the host (``seed`` for reproducibility), executed as batched kernels."""

    def __init__(self, n: int, m: int, seed: Optional[int] = None):
        self.n, self.m = n, m
        self.rng = _random.Random(seed)
        self.np_rng = np.random.RandomState(seed)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        x = _check(x)
        N, _, H, W = x.shape
        r, nr = self.rng, self.np_rng
        for _ in range(self.n):
            ops = [r.choice(_POOL) for _ in range(N)]
            vals = []
            for _, lo, hi in ops:
                v = nr.randint(1, self.m)
                vals.append(float(v) / 30 * float(hi - lo) + lo)
            for name in dict.fromkeys(o[0] for o in ops):
                idx = [i for i, o in enumerate(ops) if o[0] == name]
                x = getattr(self, "_" + name)(x, idx, [vals[i] for i in idx], N, H, W)
        return x

    # each operation: sample per-image parameters for the selected images, one launch for the batch
    def _mask_of(self, idx, fired, N):
        m = np.zeros(N, np.int32)
        for i, f in zip(idx, fired):
            m[i] = 1 if f else 0
        return m

    def _HSV(self, x, idx, vals, N, H, W):                       # :51-57
        dh, ds, dv, fired = np.zeros(N), np.zeros(N), np.zeros(N), []
        for i, v in zip(idx, vals):
            v = -v if self.rng.random() < 0.5 else v
            fired.append(self.rng.random() < 0.5)
            dh[i], ds[i], dv[i] = (np.rint(self.rng.uniform(-v, v)) for _ in range(3))   # cv2.add rounds
        return hsv_shift(x, dh, ds, dv, self._mask_of(idx, fired, N))

    def _Noise(self, x, idx, vals, N, H, W):                     # :59-63
        scale, fired = torch.zeros(N), []
        for i, v in zip(idx, vals):
            fired.append(self.rng.random() < 0.5)
            scale[i] = self.rng.uniform(0.0, v * 255)
        noise = torch.randn(N, 1, H, W, device=x.device) * scale.to(x.device).view(N, 1, 1, 1)
        return add_noise(x, noise, self._mask_of(idx, fired, N))

    def _Scale_Resize_Crop(self, x, idx, vals, N, H, W):         # :65-70
        # RandomScale (p 0.5) -> Resize(S+20) -> RandomCrop(S): two bicubic resamplings like the
        # reference's, batched by grouping images on the intermediate size
        out = x.clone()
        groups = {}
        for i, v in zip(idx, vals):
            s = self.rng.uniform(1 - v, 1 + v) if self.rng.random() < 0.5 else 1.0
            hs, ws = max(int(round(H * s)), 1), max(int(round(W * s)), 1)
            groups.setdefault((hs, ws), []).append(i)
        for (hs, ws), members in groups.items():
            sel = torch.as_tensor(members, device=x.device)
            sub = x.index_select(0, sel)
            if (hs, ws) != (H, W):
                sub = warp_affine(sub, np.tile(resize_matrix_inv(H, W, hs, ws), (len(members), 1)), (hs, ws),
                                  clamp_border=True)
            sub = warp_affine(sub, np.tile(resize_matrix_inv(hs, ws, H + 20, W + 20), (len(members), 1)),
                              (H + 20, W + 20), clamp_border=True)
            top = [self.rng.randint(0, 20) for _ in members]
            left = [self.rng.randint(0, 20) for _ in members]
            out.index_copy_(0, sel, flip_crop(sub, top, left, [0] * len(members), (H, W)))
        return out

    def _Shift_Scale_Rotate(self, x, idx, vals, N, H, W):        # :72-79
        minv, fired = np.tile(np.array([1, 0, 0, 0, 1, 0], np.float32), (N, 1)), []
        for i, v in zip(idx, vals):
            v = -v if self.rng.random() < 0.5 else v
            fired.append(self.rng.random() < 0.5)
            angle = self.rng.uniform(-90, 90)
            scale = self.rng.uniform(1 - (v + 0.5), 1 + (v + 0.5))
            dx, dy = self.rng.uniform(-v, v), self.rng.uniform(-v, v)
            minv[i] = rotation_matrix_inv(W / 2, H / 2, angle, scale, dx * W, dy * H)
        return warp_affine(x, minv, None, self._mask_of(idx, fired, N))

    def _Color(self, x, idx, vals, N, H, W):                     # :81-83
        delta = np.zeros((N, 3), np.float32)
        for i in idx:
            delta[i] = [self.rng.normalvariate(0, self.rng.uniform(-0.035, 0.035)) for _ in range(3)]
        return hed_jitter(x, delta, self._mask_of(idx, [True] * len(idx), N))

    def _Blur_img(self, x, idx, vals, N, H, W):                  # :85-89
        k, fired = np.ones(N, np.int32), []
        for i, v in zip(idx, vals):
            fired.append(self.rng.random() < 0.5)
            k[i] = self.rng.choice(list(range(3, int(v + 5) + 1, 2)))
        return box_blur(x, k, self._mask_of(idx, fired, N))

    def _bc(self, x, idx, N, b_lim, c_lim):
        alpha, beta, fired = np.ones(N, np.float32), np.zeros(N, np.float32), []
        for i, (bl, cl) in zip(idx, zip(b_lim, c_lim)):
            fired.append(self.rng.random() < 0.5)
            alpha[i] = 1.0 + self.rng.uniform(-cl, cl)
            beta[i] = self.rng.uniform(-bl, bl)
        return brightness_contrast(x, alpha, beta, self._mask_of(idx, fired, N))

    def _Brightness(self, x, idx, vals, N, H, W):                # :91-95
        return self._bc(x, idx, N, vals, [0.2] * len(idx))

    def _Contrast(self, x, idx, vals, N, H, W):                  # :97-101
        return self._bc(x, idx, N, [0.2] * len(idx), vals)

    def _Rotate_Crop(self, x, idx, vals, N, H, W):               # :103-110
        minv, fired = np.tile(np.array([1, 0, 0, 0, 1, 0], np.float32), (N, 1)), []
        for i, v in zip(idx, vals):
            v = -v if self.rng.random() < 0.5 else v
            fh = fv = False
            if self.rng.random() < 0.5:                          # Flip(): d in {-1, 0, 1}
                d = self.rng.randint(-1, 1)
                fh, fv = d in (-1, 1), d in (-1, 0)
            angle = self.rng.uniform(-v, v) if self.rng.random() < 0.5 else 0.0
            fired.append(fh or fv or angle != 0.0)
            minv[i] = rotation_matrix_inv(W / 2, H / 2, angle, 1.0, 0, 0, fh, fv, W, H)
        return warp_affine(x, minv, None, self._mask_of(idx, fired, N))


class TransformFix:
    """dataset.py:663-677 over a batch: ``weak, strong = TransformFix(image_size, N)(x)``."""

    def __init__(self, image_size: int, N: int, seed: Optional[int] = None):
        self.size = image_size
        self.rng = _random.Random(seed)
        self.randaugment = RandAugment(n=N, m=10, seed=seed)

    def _flip_crop(self, x):
        n, _, Hs, Ws = x.shape
        S = self.size
        flip = [self.rng.random() < 0.5 for _ in range(n)]
        top = [self.rng.randint(0, Hs - S) for _ in range(n)]
        left = [self.rng.randint(0, Ws - S) for _ in range(n)]
        return flip_crop(x, top, left, flip, (S, S))

    def __call__(self, x: torch.Tensor):
        x = _check(x)
        return self._flip_crop(x), self.randaugment(self._flip_crop(x))
