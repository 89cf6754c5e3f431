"""Weak / strong augmentation.

CPU part (-m "not gpu"): the numpy oracle (oracle/ref_augment.py) is pinned against the one real
dependency of the reference's augmentation stack that is installed here, OpenCV -- cv2.blur,
cv2.cvtColor RGB<->HSV, cv2.warpAffine / cv2.resize (INTER_CUBIC, BORDER_REFLECT_101) -- and against
closed forms for the rest.  GPU part: every kernel through the C ABI against the oracle on the same
seeded batch and per-image parameters -- bit-exact for the integer / table / single-rounding
operations, one grey level for the bicubic warp (fp32 vs fp64 accumulation).
"""
import numpy as np
import pytest
import torch

from oracle import ref_augment as R


def batch(n, h, w, seed):
    rng = np.random.RandomState(seed)
    x = rng.randint(0, 256, (n, 3, h, w)).astype(np.uint8)
    # smooth, tissue-like content in half of the images (interpolation tests are meaningless on noise)
    yy, xx = np.mgrid[0:h, 0:w]
    for i in range(0, n, 2):
        for c in range(3):
            x[i, c] = (127 + 100 * np.sin(xx / (5.0 + c) + i) * np.cos(yy / (7.0 + i))).astype(np.uint8)
    return x


# ------------------------------------------------------------------------------ CPU: oracle vs OpenCV
def test_oracle_blur_equals_cv2_blur():
    cv2 = pytest.importorskip("cv2")
    x = batch(3, 37, 41, 1)
    for k in (3, 5, 7):
        got = R.box_blur(x, [k] * 3)
        for n in range(3):
            ref = cv2.blur(np.ascontiguousarray(np.transpose(x[n], (1, 2, 0))), (k, k))
            assert np.array_equal(np.transpose(got[n], (1, 2, 0)), ref), k


def test_oracle_hsv_conversions_equal_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(2)
    rgb = rng.randint(0, 256, (200, 200, 3)).astype(np.uint8)
    hsv = cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV)
    assert np.array_equal(R.rgb2hsv_u8(rgb), hsv)                       # integer path: exact
    back = cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB).astype(int)
    mine = R.hsv2rgb_u8(hsv).astype(int)
    # float path: cv2 4.13 lands one level higher where the exact value is an integer (round trips
    # of real RGB triples) -- never more than one level, on < 2 % of the values
    assert np.abs(mine - back).max() <= 1 and (mine != back).mean() < 0.02
    # shift_hsv end to end against the albumentations 0.1.8 recipe run on cv2
    x = np.transpose(rgb, (2, 0, 1))[None]
    got = R.hsv_shift(x, [7], [-20], [13])[0]
    h, s, v = cv2.split(hsv.astype(np.int32))
    h = h + 7
    h = np.where(h < 0, h + 180, h); h = np.where(h > 180, h - 180, h)
    ref = cv2.cvtColor(cv2.merge((h.astype(np.uint8), np.clip(s - 20, 0, 255).astype(np.uint8),
                                  np.clip(v + 13, 0, 255).astype(np.uint8))), cv2.COLOR_HSV2RGB)
    d = np.abs(np.transpose(got, (1, 2, 0)).astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.02


def test_oracle_bicubic_warp_tracks_cv2_warpaffine_and_resize():
    """cv2 quantises coordinates to 1/32 pixel and uses fixed-point weights; the oracle (and the
    kernel) interpolate at the exact coordinate: a few grey levels on smooth content."""
    cv2 = pytest.importorskip("cv2")
    x = batch(2, 64, 64, 3)[:1]                                         # the smooth image
    img = np.ascontiguousarray(np.transpose(x[0], (1, 2, 0)))
    M = cv2.getRotationMatrix2D((32, 32), 23.0, 1.1)
    M[0, 2] += 2.5; M[1, 2] -= 1.25
    ref = cv2.warpAffine(img, M, (64, 64), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REFLECT_101)
    got = R.warp_affine(x, [R.rotation_matrix_inv(32, 32, 23.0, 1.1, 2.5, -1.25)], 64, 64)[0]
    d = np.abs(np.transpose(got, (1, 2, 0)).astype(int) - ref.astype(int))
    assert d.max() <= 6 and d.mean() < 0.6, (d.max(), d.mean())
    ref = cv2.resize(img, (84, 84), interpolation=cv2.INTER_CUBIC)
    got = R.warp_affine(x, [R.resize_matrix_inv(64, 64, 84, 84)], 84, 84, clamp_border=True)[0]
    d = np.abs(np.transpose(got, (1, 2, 0)).astype(int) - ref.astype(int))
    assert d.max() <= 2 and d.mean() < 0.3, (d.max(), d.mean())
    # identity and pure flips are exact pixel permutations
    assert np.array_equal(R.warp_affine(x, [[1, 0, 0, 0, 1, 0]], 64, 64), x)
    from ssl_cr_histo_b200.augment import rotation_matrix_inv
    flipped = R.warp_affine(x, [rotation_matrix_inv(32, 32, 0.0, 1.0, 0, 0, True, False, 64, 64)], 64, 64)
    assert np.array_equal(flipped, x[:, :, :, ::-1])


def test_oracle_closed_forms():
    x = batch(4, 20, 24, 4)
    # flip + crop
    got = R.flip_crop(x, [1, 0, 3, 2], [2, 4, 0, 1], [0, 1, 1, 0], 16, 18)
    assert np.array_equal(got[0], x[0, :, 1:17, 2:20])
    assert np.array_equal(got[1], x[1, :, :, ::-1][:, 0:16, 4:22])
    # brightness / contrast: identity parameters, saturation, truncation
    assert np.array_equal(R.brightness_contrast(x, [1] * 4, [0] * 4), x)
    assert R.brightness_contrast(x, [2.0] * 4, [40.0] * 4).max() == 255
    assert np.array_equal(R.brightness_contrast(x, [0.5] * 4, [0] * 4), x // 2)
    # H&E-DAB jitter: zero offsets reproduce the image up to the truncating cast
    same = R.hed_jitter(x, np.zeros((4, 3)))
    assert np.abs(same.astype(int) - x.astype(int)).max() <= 1
    darker = R.hed_jitter(x, np.full((4, 3), 0.03))
    assert darker.astype(int).mean() < x.astype(int).mean()
    # noise: zero noise is the identity; apply mask copies through
    assert np.array_equal(R.add_noise(x, np.zeros((4, 1, 20, 24))), x)
    assert np.array_equal(R.box_blur(x, [5] * 4, apply=[0, 0, 0, 0]), x)


def test_randaugment_pool_and_magnitudes_follow_the_reference():
    from ssl_cr_histo_b200 import augment
    assert [p[0] for p in augment._POOL] == ["HSV", "Noise", "Scale_Resize_Crop", "Shift_Scale_Rotate",
                                             "Color", "Blur_img", "Brightness", "Contrast", "Rotate_Crop"]
    assert [(p[1], p[2]) for p in augment._POOL] == [(-1, 1), (0, 0.15), (0.8, 1.2), (0.01, 0.1),
                                                     (-0.035, 0.035), (0, 2), (-0.2, 0.2), (-0.2, 0.2),
                                                     (-90, 90)]                 # models/randaugment.py:112-123
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        augment.flip_crop(torch.zeros(1, 3, 8, 8, dtype=torch.uint8), [0], [0], [0], (8, 8))


# ------------------------------------------------------------------------------ GPU: kernels vs oracle
gpu = pytest.mark.gpu
DEV = "cuda"


@gpu
@pytest.mark.parametrize("shape", [(5, 64, 64), (3, 37, 53), (16, 224, 224)])
def test_pointwise_kernels_bit_exact(shape):
    from ssl_cr_histo_b200 import augment
    n, h, w = shape
    x = batch(n, h + 9, w + 6, 10)
    xd = torch.from_numpy(x).to(DEV)
    rng = np.random.RandomState(11)
    top, left, flip = rng.randint(0, 10, n), rng.randint(0, 7, n), rng.randint(0, 2, n)
    got = augment.flip_crop(xd, top, left, flip, (h, w))
    ref = R.flip_crop(x, top, left, flip, h, w)
    assert np.array_equal(got.cpu().numpy(), ref)
    x, xd = ref, got
    apply = rng.randint(0, 2, n); apply[0] = 1
    alpha = (1 + rng.uniform(-0.2, 0.2, n)).astype(np.float32)
    beta = rng.uniform(-0.2, 0.2, n).astype(np.float32)
    for by_max in (False, True):
        off = beta * (255.0 if by_max else R.image_mean(x))
        ref = R.brightness_contrast(x, alpha, off.astype(np.float32), apply)
        got = augment.brightness_contrast(xd, alpha, beta, apply, beta_by_max=by_max)
        assert np.array_equal(got.cpu().numpy(), ref), by_max
    dh, ds, dv = rng.randint(-30, 31, n), rng.randint(-40, 41, n), rng.randint(-40, 41, n)
    assert np.array_equal(augment.hsv_shift(xd, dh, ds, dv, apply).cpu().numpy(), R.hsv_shift(x, dh, ds, dv, apply))
    assert np.array_equal(augment.hsv_shift(xd, [0] * n, [0] * n, [0] * n).cpu().numpy(),
                          R.hsv_shift(x, [0] * n, [0] * n, [0] * n))
    noise = (rng.randn(n, 1, h, w) * rng.uniform(0, 38, (n, 1, 1, 1))).astype(np.float32)
    assert np.array_equal(augment.add_noise(xd, noise, apply).cpu().numpy(), R.add_noise(x, noise, apply))
    ks = rng.choice([1, 3, 5, 7], n)
    assert np.array_equal(augment.box_blur(xd, ks, apply).cpu().numpy(), R.box_blur(x, ks, apply))
    delta = rng.uniform(-0.035, 0.035, (n, 3)).astype(np.float32)
    got, ref = augment.hed_jitter(xd, delta, apply).cpu().numpy().astype(int), R.hed_jitter(x, delta, apply).astype(int)
    # double precision on both sides; libm vs CUDA exp/log differ in the last ulp, which can move a
    # truncation boundary (and 0 <-> 255 where a negative value wraps, as in the reference)
    assert (got != ref).mean() < 1e-5


@gpu
@pytest.mark.parametrize("shape", [(4, 64, 64), (3, 50, 70)])
def test_bicubic_warp_kernel(shape):
    from ssl_cr_histo_b200 import augment
    n, h, w = shape
    x = batch(n, h, w, 20)
    xd = torch.from_numpy(x).to(DEV)
    rng = np.random.RandomState(21)
    minv = np.stack([augment.rotation_matrix_inv(w / 2, h / 2, rng.uniform(-90, 90), rng.uniform(0.5, 1.6),
                                                 rng.uniform(-5, 5), rng.uniform(-5, 5), i % 2 == 0, i % 3 == 0, w, h)
                     for i in range(n)])
    apply = np.ones(n, np.int32); apply[-1] = 0
    got = augment.warp_affine(xd, minv, None, apply).cpu().numpy().astype(int)
    ref = R.warp_affine(x, minv, h, w, apply).astype(int)
    assert np.abs(got - ref).max() <= 1 and (got != ref).mean() < 2e-3     # fp32 vs fp64 accumulation
    assert np.array_equal(got[-1], x[-1])                                   # copied through
    up = augment.warp_affine(xd, np.tile(augment.resize_matrix_inv(h, w, h + 20, w + 20), (n, 1)),
                             (h + 20, w + 20), clamp_border=True).cpu().numpy().astype(int)
    ref = R.warp_affine(x, np.tile(R.resize_matrix_inv(h, w, h + 20, w + 20), (n, 1)), h + 20, w + 20,
                        clamp_border=True).astype(int)
    assert np.abs(up - ref).max() <= 1 and (up != ref).mean() < 2e-3
    ident = augment.warp_affine(xd, np.tile(np.array([1, 0, 0, 0, 1, 0], np.float32), (n, 1)))
    assert torch.equal(ident, xd)


@gpu
def test_transformfix_views_feed_the_trunk():
    """TransformFix over a batch (dataset.py:663-677): shapes / dtypes, determinism under a seed, the
    weak view is a flip+crop of the input, the strong view differs, and both go straight into the
    uint8 input path of the drop-in model."""
    from ssl_cr_histo_b200 import augment
    import ssl_cr_histo_b200.net as net
    x = torch.from_numpy(batch(12, 72, 72, 30)).to(DEV)
    tf = augment.TransformFix(64, 7, seed=5)
    weak, strong = tf(x)
    weak2, strong2 = augment.TransformFix(64, 7, seed=5)(x)
    assert weak.shape == strong.shape == (12, 3, 64, 64) and weak.dtype == strong.dtype == torch.uint8
    assert torch.equal(weak, weak2)
    assert not torch.equal(weak, strong)
    # every weak image is one of the 2 * 9 * 9 flip/crop windows of its source
    src = x[0].cpu().numpy()
    found = any(np.array_equal(weak[0].cpu().numpy(), (src[:, :, ::-1] if f else src)[:, t:t + 64, l:l + 64])
                for f in (0, 1) for t in range(9) for l in range(9))
    assert found
    torch.manual_seed(0)
    model = net.TripletNet_Finetune("resnet18").to(DEV).eval()
    with torch.no_grad():
        f = model(torch.cat((weak, strong)))
    assert f.shape == (24, 768) and torch.isfinite(f).all()
