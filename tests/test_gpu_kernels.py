"""GPU parity tests, kernel by kernel, through the C ABI (libb2n.so via ctypes) against plain
torch fp32 CPU references of the same op.  Integer-valued inputs make TF32 products and FP32
sums exact, so the tensor-core kernels are required to be BIT-EXACT there; element-wise and
FP32-GEMM kernels are held to fp32 round-off."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from ssl_cr_histo_b200 import _lib, losses, weights
from ssl_cr_histo_b200._lib import call
from oracle import ref_net as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ints(shape, lo, hi, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(lo, hi + 1, shape, generator=g).float() * scale


def to_nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def from_nhwc(t):
    return t.permute(0, 3, 1, 2).contiguous()


def pack_fwd(w, split=False):
    """split: the library's (hi, lo) FP16 pack; plain: the same K-major layout in fp32 (built with
    torch -- the values used with it are TF32-exact)."""
    K, C, R, S = w.shape
    if not split:
        return w.permute(0, 2, 3, 1).reshape(K, R * S * C).contiguous().to(DEV)
    out = torch.empty(2, K, R * S * C, device=DEV, dtype=torch.float16)
    call("b2n_pack_weight_fwd", w.to(DEV), out[0], out[1], K, C, R, S)
    return out[0], out[1]


def split_pair(t):
    """(hi, lo) FP16 pair of an fp32 tensor, as the producing kernels store activations."""
    hi = t.clamp(-65504, 65504).half()
    return hi, (t - hi.float()).half()


def pack_dgrad(w):
    K, C, R, S = w.shape
    out = torch.empty(C, R * S * K, device=DEV)
    call("b2n_pack_weight_dgrad", w.to(DEV), out, K, C, R, S)
    return out


def conv(x_nhwc, wp, N, H, W, Cin, Cout, R, stride, plo, phi, **kw):
    """x_nhwc / wp: plain fp32 tensors (one TF32 pass) or (hi, lo) FP16 pairs (compensated)."""
    P = (H + plo + phi - R) // stride + 1
    Q = (W + plo + phi - R) // stride + 1
    y = torch.full((N, P, Q, Cout), float("nan"), device=DEV)
    if isinstance(x_nhwc, tuple):
        x32, (xh, xl), w32, (wh, wl) = None, x_nhwc, None, wp
    else:
        x32, xh, xl, w32, wh, wl = x_nhwc, None, None, wp, None, None
    yp, rp = kw.get("y_pair", (None, None)), kw.get("resid_pair", (None, None))
    call("b2n_conv_fwd", x32, xh, xl, w32, wh, wl, y, yp[0], yp[1], N, H, W, Cin, Cout, R, R, stride,
         plo, phi, plo, phi, kw.get("scale"), kw.get("shift"), kw.get("resid"), rp[0], rp[1],
         kw.get("mask"), kw.get("relu", 0), kw.get("rnd", 0), kw.get("stats"), kw.get("lo_flag"),
         0, 0, 0, 0, 0, kw.get("gate"), *kw.get("bnb", (None,) * 5))
    return y


def test_device_is_blackwell():
    assert _lib.load().b2n_device_ok() == 1, "tests need a compute-capability 10.x device"


CONV_CASES = [  # N, H, W, Cin, Cout, R, stride, pad   (every ResNet18 conv shape class + ragged M)
    (2, 56, 56, 64, 64, 3, 1, 1),
    (2, 56, 56, 64, 128, 3, 2, 1),
    (2, 56, 56, 64, 128, 1, 2, 0),
    (3, 28, 28, 128, 128, 3, 1, 1),
    (2, 28, 28, 128, 256, 3, 2, 1),
    (3, 14, 14, 256, 256, 3, 1, 1),
    (2, 14, 14, 256, 512, 1, 2, 0),
    (5, 7, 7, 512, 512, 3, 1, 1),
    (1, 7, 7, 256, 512, 3, 2, 1),       # odd input, 4x4 output
    (1, 4, 4, 64, 64, 3, 1, 1),         # M = 16 < one tile
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_bit_exact_with_bn_statistics(case):
    N, H, W, Cin, Cout, R, s, p = case
    x = ints((N, Cin, H, W), -4, 4, 1)
    w = ints((Cout, Cin, R, R), -2, 2, 2, 0.25)
    ref = F.conv2d(x, w, None, s, p)
    stats = torch.zeros(2 * Cout, device=DEV, dtype=torch.float64)
    y = conv(to_nhwc(x).to(DEV), pack_fwd(w), N, H, W, Cin, Cout, R, s, p, p, stats=stats)
    assert torch.equal(from_nhwc(y.cpu()), ref)                          # one TF32 pass
    assert torch.allclose(stats[:Cout].cpu(), ref.double().sum((0, 2, 3)), rtol=1e-12, atol=1e-6)
    assert torch.allclose(stats[Cout:].cpu(), ref.double().pow(2).sum((0, 2, 3)), rtol=1e-6)
    if Cin % 64 == 0:                                                    # FP16 (hi, lo) mode
        xh, xl = split_pair(to_nhwc(x))
        assert float(xl.abs().max()) == 0
        stats.zero_()
        y = conv((xh.to(DEV), xl.to(DEV)), pack_fwd(w, split=True), N, H, W, Cin, Cout, R, s, p, p,
                 stats=stats)
        assert torch.equal(from_nhwc(y.cpu()), ref)
        assert torch.allclose(stats[:Cout].cpu(), ref.double().sum((0, 2, 3)), rtol=1e-12, atol=1e-6)


@pytest.mark.parametrize("case", [CONV_CASES[0], CONV_CASES[1], CONV_CASES[7], (2, 32, 32, 32, 64, 4, 1, None)])
def test_conv_split_mode_is_fp32_accurate(case):
    """Error-compensated mode: generic fp32 operands as (hi, lo) FP16 pairs -> result at the
    tensor core's FP32-accumulation floor, where a single TF32 pass is only good to ~5e-4; also
    checks the (hi, lo) output path and the FP16-pair residual."""
    N, H, W, Cin, Cout, R, s, p = case
    plo, phi = (p, p) if p is not None else (2, 1)            # the stem's asymmetric padding
    g = torch.Generator().manual_seed(77)
    x = torch.randn(N, Cin, H, W, generator=g) * 3 + 1
    w = torch.randn(Cout, Cin, R, R, generator=g) * 0.05
    xp = F.pad(x.double(), (plo, phi, plo, phi))
    ref = F.conv2d(xp, w.double(), None, s, 0)
    scale = float(ref.abs().max())
    xh, xl = split_pair(to_nhwc(x))
    y = conv((xh.to(DEV), xl.to(DEV)), pack_fwd(w, split=True), N, H, W, Cin, Cout, R, s, plo, phi)
    err3 = float((from_nhwc(y.cpu()).double() - ref).abs().max()) / scale
    y1 = conv(O.tf32_round(to_nhwc(x)).to(DEV), O.tf32_round(pack_fwd(w).cpu()).to(DEV), N, H, W, Cin,
              Cout, R, s, plo, phi)
    err1 = float((from_nhwc(y1.cpu()).double() - ref).abs().max()) / scale
    assert err3 < 1e-4, err3      # floor = the tensor core's FP32 accumulation (K up to 4608)
    assert err1 > 4 * err3                                     # the compensation is what does it
    res = torch.randn(ref.shape, generator=g)
    rh, rl = split_pair(to_nhwc(res))
    y_h, y_l = torch.empty_like(y, dtype=torch.half), torch.empty_like(y, dtype=torch.half)
    y2 = conv((xh.to(DEV), xl.to(DEV)), pack_fwd(w, split=True), N, H, W, Cin, Cout, R, s, plo, phi,
              resid_pair=(rh.to(DEV), rl.to(DEV)), relu=1, y_pair=(y_h, y_l))
    want = torch.relu(ref + res.double())
    got = from_nhwc((y_h.double() + y_l.double()).cpu())
    assert float((got - want).abs().max()) / scale < 1e-4
    assert float((from_nhwc(y2.cpu()).double() - want).abs().max()) / scale < 1e-4   # fp32 copy


def test_conv_epilogue_scale_shift_resid_mask_relu_round():
    N, H, W, Cin, Cout = 2, 14, 14, 64, 128
    x, w = ints((N, Cin, H, W), -4, 4, 3), ints((Cout, Cin, 3, 3), -2, 2, 4, 0.25)
    scale, shift = ints((Cout,), 1, 4, 5, 0.5), ints((Cout,), -8, 8, 6)
    resid, mask = ints((N, Cout, H, W), -16, 16, 7), ints((N, Cout, H, W), -1, 1, 8)
    acc = F.conv2d(x, w, None, 1, 1)
    ref = torch.relu(acc * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
                     + torch.where(mask > 0, resid, torch.zeros(())))
    y = conv(to_nhwc(x).to(DEV), pack_fwd(w), N, H, W, Cin, Cout, 3, 1, 1, 1, scale=scale.to(DEV),
             shift=shift.to(DEV), resid=to_nhwc(resid).to(DEV), mask=to_nhwc(mask).to(DEV), relu=1,
             rnd=1)
    assert torch.equal(from_nhwc(y.cpu()), O.tf32_round(ref))


@pytest.mark.parametrize("shape", [(3, 56, 56, 64, 64), (2, 28, 28, 128, 128), (5, 7, 7, 512, 512),
                                   (1, 5, 9, 64, 64)])
def test_conv_dgrad_epilogue_gate_and_batchnorm_backward_sums(shape):
    """What the trunk's backward launches: (a) conv1's data gradient + the (pre-gated) shortcut
    gradient, zeroed by the ReLU gate of the block input; (b) conv2's data gradient with bn1's
    ReLU gate recomputed from y and the two BatchNorm-backward sums taken in the epilogue --
    against b2n_bn_bwd_reduce over the materialised tensors (integer-valued data: exact)."""
    N, H, W, Cin, Cout = shape
    w = ints((Cout, Cin, 3, 3), -2, 2, 51, 0.25)
    dy = ints((N, Cout, H, W), -4, 4, 52)
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), w, dy, stride=1, padding=1)
    resid, act = ints((N, Cin, H, W), -16, 16, 53), ints((N, Cin, H, W), -1, 1, 54)
    g = to_nhwc(dy).to(DEV)
    out = conv(g, pack_dgrad(w), N, H, W, Cout, Cin, 3, 1, 1, 1, resid=to_nhwc(resid).to(DEV),
               gate=to_nhwc(act).to(DEV))
    assert torch.equal(from_nhwc(out.cpu()), torch.where(act > 0, ref + resid, torch.zeros(())))
    out = conv(g, pack_dgrad(w), N, H, W, Cout, Cin, 3, 1, 1, 1, gate=to_nhwc(act).to(DEV))
    assert torch.equal(from_nhwc(out.cpu()), torch.where(act > 0, ref, torch.zeros(())))
    # (b) y: raw output of the BN in front of the ReLU; small integers keep every product exact
    y = ints((N, Cin, H, W), -3, 3, 55)
    mean, invstd = ints((Cin,), -1, 1, 56, 0.5), ints((Cin,), 1, 2, 57, 0.5)
    scale, shift = ints((Cin,), 1, 2, 58), ints((Cin,), -2, 2, 59, 0.5)
    yd = to_nhwc(y).to(DEV)
    stats = torch.zeros(2 * Cin, device=DEV, dtype=torch.float64)
    out = conv(g, pack_dgrad(w), N, H, W, Cout, Cin, 3, 1, 1, 1, stats=stats,
               bnb=(yd, mean.to(DEV), invstd.to(DEV), scale.to(DEV), shift.to(DEV)))
    gate = (y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)) > 0
    want = torch.where(gate, ref, torch.zeros(()))
    assert torch.equal(from_nhwc(out.cpu()), want)
    xhat = (y - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
    # (per-CTA partial sums are fp32: exact up to the rounding of long sums)
    tol = dict(rtol=1e-6, atol=1e-5 * float((want.abs() * xhat.abs()).sum((0, 2, 3)).max()))
    assert torch.allclose(stats[:Cin].cpu(), want.double().sum((0, 2, 3)), **tol)
    assert torch.allclose(stats[Cin:].cpu(), (want.double() * xhat.double()).sum((0, 2, 3)), **tol)
    s_ref = torch.zeros(2 * Cin, device=DEV, dtype=torch.float64)
    call("b2n_bn_bwd_reduce", to_nhwc(ref).to(DEV), None, yd, mean.to(DEV), invstd.to(DEV),
         scale.to(DEV), shift.to(DEV), s_ref, N * H * W, Cin)
    assert torch.allclose(stats, s_ref, **tol)
    # without the y-derived gate: plain sums of the result
    stats.zero_()
    out = conv(g, pack_dgrad(w), N, H, W, Cout, Cin, 3, 1, 1, 1, stats=stats,
               bnb=(yd, mean.to(DEV), invstd.to(DEV), None, None))
    assert torch.equal(from_nhwc(out.cpu()), ref)
    assert torch.allclose(stats[:Cin].cpu(), ref.double().sum((0, 2, 3)), **tol)
    with pytest.raises(RuntimeError, match="exclusive"):
        conv(g, pack_dgrad(w), N, H, W, Cout, Cin, 3, 1, 1, 1, resid=to_nhwc(resid).to(DEV),
             mask=to_nhwc(act).to(DEV), gate=to_nhwc(act).to(DEV))


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_wgrad_bit_exact(case):
    N, H, W, Cin, Cout, R, s, p = case
    x = ints((N, Cin, H, W), -2, 2, 11)
    P = (H + 2 * p - R) // s + 1
    dy = ints((N, Cout, P, P), -2, 2, 12, 0.5)
    ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, R, R), dy, stride=s, padding=p)
    dwp = torch.zeros(Cout, R * R * Cin, device=DEV)
    call("b2n_conv_wgrad", to_nhwc(x).to(DEV), to_nhwc(dy).to(DEV), dwp, N, H, W, Cin, Cout, R, R, s,
         p, p, p, p, 0, 0)
    dw = torch.empty(Cout, Cin, R, R, device=DEV)
    call("b2n_unpack_wgrad", dwp, dw, Cout, Cin, R, R, 0, 1)
    assert torch.equal(dw.cpu(), ref)
    call("b2n_unpack_wgrad", dwp, dw, Cout, Cin, R, R, 1, 1)   # accumulate: a second writer of the slot
    assert torch.equal(dw.cpu(), 2 * ref)
    # deterministic mode: one plane per split-K group, no atomics, no zero-fill, summed by the unpack
    planes = _lib.wgrad_planes(N, H, W, Cin, Cout, R, R, s, p, p, p, p)
    part = torch.full((planes, Cout, R * R * Cin), float("nan"), device=DEV)
    call("b2n_conv_wgrad", to_nhwc(x).to(DEV), to_nhwc(dy).to(DEV), part, N, H, W, Cin, Cout, R, R, s,
         p, p, p, p, 1, 0)
    call("b2n_unpack_wgrad", part, dw, Cout, Cin, R, R, 0, planes)
    assert torch.equal(dw.cpu(), ref)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_dgrad_bit_exact(case):
    """data gradient = forward kernel over the (zero-stuffed) output gradient with the flipped,
    transposed weight pack."""
    N, H, W, Cin, Cout, R, s, p = case
    w = ints((Cout, Cin, R, R), -2, 2, 21, 0.25)
    P = (H + 2 * p - R) // s + 1
    dy = ints((N, Cout, P, P), -4, 4, 22)
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), w, dy, stride=s, padding=p)
    g = to_nhwc(dy).to(DEV)
    if s == 2:
        up = torch.full((N, H, W, Cout), float("nan"), device=DEV)
        call("b2n_upsample_zero", g, up, N, P, P, H, W, Cout)
        g, gh = up, H
    else:
        gh = P
    pad = R - 1 - p
    dx = conv(g, pack_dgrad(w), N, gh, gh, Cout, Cin, R, 1, pad, pad)
    assert torch.equal(from_nhwc(dx.cpu()), ref)


@pytest.mark.parametrize("shape", [(2, 56, 56, 64, 128), (3, 7, 7, 256, 512), (2, 14, 10, 128, 256)])
def test_conv_dgrad_stride2_parity_classes_bit_exact(shape):
    """stride-2 3x3 data gradient as four parity-class convs + the 1x1 shortcut gradient on the
    even pixels, accumulated in place -- exactly what the trunk's backward pass launches."""
    N, H, W, Cin, Cout = shape
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    w3, w1 = ints((Cout, Cin, 3, 3), -2, 2, 41, 0.25), ints((Cout, Cin, 1, 1), -2, 2, 42, 0.25)
    dy3, dy1 = ints((N, Cout, P, Q), -4, 4, 43), ints((N, Cout, P, Q), -4, 4, 44)
    ref = (torch.nn.grad.conv2d_input((N, Cin, H, W), w3, dy3, stride=2, padding=1)
           + torch.nn.grad.conv2d_input((N, Cin, H, W), w1, dy1, stride=2, padding=0))
    g_in = torch.full((N, H, W, Cin), float("nan"), device=DEV)
    buf = torch.empty(9 * Cin * Cout, device=DEV)
    call("b2n_pack_weight_dgrad_s2", w3.to(DEV), buf, Cout, Cin)
    act = ints((N, Cin, H, W), -1, 1, 45)            # the block input whose ReLU gate the classes apply
    gate_t = to_nhwc(act).to(DEV)

    def launch(x, wp, R, S, phi_h, phi_w, resid, place, gate=None):
        call("b2n_conv_fwd", x, None, None, wp, None, None, g_in, None, None, N, P, Q, Cout, Cin, R, S,
             1, 0, phi_h, 0, phi_w, None, None, resid, None, None, None, 0, 0, None, None, *place,
             gate, None, None, None, None, None)

    launch(to_nhwc(dy1).to(DEV), pack_dgrad(w1), 1, 1, 0, 0, None, (2, 0, 0, H, W))
    d3, off = to_nhwc(dy3).to(DEV), 0
    for cls, (a0, b0) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        n = Cin * (1 + a0) * (1 + b0) * Cout
        launch(d3, buf[off:off + n], 1 + a0, 1 + b0, a0, b0, g_in if cls == 0 else None,
               (2, a0, b0, H, W))
        off += n
    assert torch.equal(from_nhwc(g_in.cpu()), ref)
    # the same with every class launch gating its quarter (gradient w.r.t. a post-ReLU input)
    g_in.fill_(float("nan"))
    launch(to_nhwc(dy1).to(DEV), pack_dgrad(w1), 1, 1, 0, 0, None, (2, 0, 0, H, W))
    off = 0
    for cls, (a0, b0) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        n = Cin * (1 + a0) * (1 + b0) * Cout
        launch(d3, buf[off:off + n], 1 + a0, 1 + b0, a0, b0, g_in if cls == 0 else None,
               (2, a0, b0, H, W), gate=gate_t)
        off += n
    assert torch.equal(from_nhwc(g_in.cpu()), torch.where(act > 0, ref, torch.zeros(())))
    # the merged launch: all four classes from one pass over dY (four TMEM accumulators per tile)
    wm = torch.empty(Cin, 9 * Cout, device=DEV)
    call("b2n_pack_weight_dgrad_s2m", w3.to(DEV), wm, Cout, Cin)
    for gate in (None, gate_t):
        g_in.fill_(float("nan"))
        launch(to_nhwc(dy1).to(DEV), pack_dgrad(w1), 1, 1, 0, 0, None, (2, 0, 0, H, W))
        call("b2n_conv_dgrad_s2", d3, wm, g_in, N, P, Q, Cout, Cin, H, W, g_in, gate)
        want = ref if gate is None else torch.where(act > 0, ref, torch.zeros(()))
        assert torch.equal(from_nhwc(g_in.cpu()), want), gate is None
    # ... with the 1x1 shortcut gradient accumulated inside the launch (a fifth K block of class (0,0))
    for gate in (None, gate_t):
        g_in.fill_(float("nan"))
        call("b2n_conv_dgrad_s2_sc", d3, wm, to_nhwc(dy1).to(DEV), pack_dgrad(w1), g_in, N, P, Q, Cout, Cin,
             H, W, gate)
        want = ref if gate is None else torch.where(act > 0, ref, torch.zeros(()))
        assert torch.equal(from_nhwc(g_in.cpu()), want), gate is None
    # without the shortcut term
    call("b2n_conv_dgrad_s2", d3, wm, g_in, N, P, Q, Cout, Cin, H, W, None, None)
    assert torch.equal(from_nhwc(g_in.cpu()),
                       torch.nn.grad.conv2d_input((N, Cin, H, W), w3, dy3, stride=2, padding=1))


@pytest.mark.parametrize("size", [(2, 64, 64), (1, 224, 224), (3, 34, 46)])
def test_stem_space_to_depth_conv_and_wgrad(size):
    N, H, W = size
    x = ints((N, 3, H, W), 0, 255, 31)                       # uint8-valued patches
    w = ints((64, 3, 7, 7), -2, 2, 32, 0.25)
    ref = F.conv2d(x, w, None, 2, 3)
    xs = torch.empty(N, H // 2, W // 2, 12, device=DEV)      # fp32 copy (wgrad): the 12 real channels
    xs_h = torch.empty(N, H // 2, W // 2, 16, device=DEV, dtype=torch.half)   # FP16 pair: 16
    xs_l = torch.empty_like(xs_h)
    flag = torch.zeros(1, device=DEV, dtype=torch.int32)
    call("b2n_stem_pack_input", x.to(DEV), xs_h, xs_l, xs, flag, N, H, W)
    assert int(flag) == 0                                     # uint8-valued input: lo plane is zero
    ws = torch.empty(2, 64, 16 * 16, device=DEV, dtype=torch.half)
    call("b2n_stem_pack_weight", w.to(DEV), ws[0], ws[1], 64)
    y = conv((xs_h, xs_l), (ws[0], ws[1]), N, H // 2, W // 2, 16, 64, 4, 1, 2, 1)
    assert torch.equal(from_nhwc(y.cpu()), ref)
    y = conv((xs_h, xs_l), (ws[0], ws[1]), N, H // 2, W // 2, 16, 64, 4, 1, 2, 1, lo_flag=flag)
    assert torch.equal(from_nhwc(y.cpu()), ref)               # lo plane skipped
    xf = x + 0.37                                             # non-integer image: flag raised
    flag.zero_()
    call("b2n_stem_pack_input", xf.to(DEV), xs_h, xs_l, None, flag, N, H, W)
    assert int(flag) == 1
    y = conv((xs_h, xs_l), (ws[0], ws[1]), N, H // 2, W // 2, 16, 64, 4, 1, 2, 1, lo_flag=flag)
    ref_f = F.conv2d(xf.double(), w.double(), None, 2, 3)
    assert float((from_nhwc(y.cpu()).double() - ref_f).abs().max()) < 1e-5 * float(ref_f.abs().max())
    dy = ints(tuple(ref.shape), -1, 1, 33)
    ref_dw = torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=2, padding=3)
    dws = torch.zeros(64, 16 * 32, device=DEV)
    call("b2n_conv_wgrad", xs, to_nhwc(dy).to(DEV), dws, N, H // 2, W // 2, 32, 64, 4, 4, 1, 2, 1, 2, 1, 0, 12)
    dw = torch.empty(64, 3, 7, 7, device=DEV)
    call("b2n_stem_unpack_wgrad", dws, dw, 64, 0, 1)
    assert torch.equal(dw.cpu(), ref_dw)
    call("b2n_stem_unpack_wgrad", dws, dw, 64, 1, 1)
    assert torch.equal(dw.cpu(), 2 * ref_dw)
    planes = _lib.wgrad_planes(N, H // 2, W // 2, 32, 64, 4, 4, 1, 2, 1, 2, 1)
    part = torch.full((planes, 64, 16 * 32), float("nan"), device=DEV)
    call("b2n_conv_wgrad", xs, to_nhwc(dy).to(DEV), part, N, H // 2, W // 2, 32, 64, 4, 4, 1, 2, 1, 2, 1, 1, 12)
    # the generic (one box per tap) kernel with the same 12-of-32 stored channels
    import os
    os.environ["B2N_NO_HALO"] = "1"
    try:
        dws.zero_()
        call("b2n_conv_wgrad", xs, to_nhwc(dy).to(DEV), dws, N, H // 2, W // 2, 32, 64, 4, 4, 1, 2, 1, 2, 1, 0, 12)
        call("b2n_stem_unpack_wgrad", dws, dw, 64, 0, 1)
    finally:
        del os.environ["B2N_NO_HALO"]
    assert torch.equal(dw.cpu(), ref_dw)
    call("b2n_stem_unpack_wgrad", part, dw, 64, 0, planes)
    assert torch.equal(dw.cpu(), ref_dw)


def test_conv_rejects_bad_shapes_loudly():
    x = torch.zeros(1, 4, 4, 24, device=DEV)
    with pytest.raises(RuntimeError, match="Cin=24"):
        conv(x, torch.zeros(64, 24 * 9, device=DEV), 1, 4, 4, 24, 64, 3, 1, 1, 1)


@pytest.mark.parametrize("C,rows,n_updates", [(64, 1000, 1), (512, 98, 3), (128, 7, 1)])
def test_batchnorm_forward_backward_and_running_stats(C, rows, n_updates):
    g = torch.Generator().manual_seed(C + rows)
    y = (torch.randn(rows, C, generator=g) * 3 + 1.5)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    res = torch.randn(rows, C, generator=g)
    gout = torch.randn(rows, C, generator=g)
    bn = torch.nn.BatchNorm1d(C)
    bn.weight.data.copy_(gamma); bn.bias.data.copy_(beta)
    bn.running_mean.data.uniform_(-1, 1, generator=g); bn.running_var.data.uniform_(0.5, 2, generator=g)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    yr = y.clone().requires_grad_(True)
    for _ in range(n_updates):
        z = bn(yr)
    out_ref = torch.relu(z + res)
    out_ref.backward(gout)

    yd = y.to(DEV)
    stats = torch.stack([yd.double().sum(0), yd.double().pow(2).sum(0)]).flatten().contiguous()
    rm, rv = rm0.to(DEV), rv0.to(DEV)
    scale, shift, mean, invstd = (torch.empty(C, device=DEV) for _ in range(4))
    inv_gamma = torch.empty(C, device=DEV)
    call("b2n_bn_finalize", stats, gamma.to(DEV), beta.to(DEV), rm, rv, scale, shift, mean, invstd,
         inv_gamma, C, float(rows), 0.1, 1e-5, n_updates)
    assert torch.allclose(inv_gamma.cpu(), 1 / gamma, rtol=1e-6)
    out = torch.empty(rows, C, device=DEV)
    call("b2n_bn_apply", yd, scale, shift, res.to(DEV), None, None, None, None, out, None, None, rows,
         C, 1, 0)
    rh, rl = split_pair(res)
    o_h = torch.empty(rows, C, device=DEV, dtype=torch.half)
    o_l, o32 = torch.empty_like(o_h), torch.empty(rows, C, device=DEV)
    call("b2n_bn_apply", yd, scale, shift, None, None, None, rh.to(DEV), rl.to(DEV), o32, o_h, o_l,
         rows, C, 1, 1)
    assert torch.allclose((o_h.float() + o_l.float()).cpu(), out_ref.detach(), rtol=1e-4, atol=1e-5)
    assert torch.equal(o32.cpu(), O.tf32_round(o32.cpu()))                  # backward copy is TF32
    assert float((o_h.float() + o_l.float() - out).abs().max()) <= 2e-6 * float(out.abs().max())
    assert torch.allclose(out.cpu(), out_ref.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(rm.cpu(), bn.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv.cpu(), bn.running_var, rtol=1e-5, atol=1e-6)
    sums = torch.zeros(2 * C, device=DEV, dtype=torch.float64)
    call("b2n_bn_bwd_reduce", gout.to(DEV), out, yd, mean, invstd, None, None, sums, rows, C)
    dy, dgamma, dbeta = torch.empty(rows, C, device=DEV), torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    call("b2n_bn_bwd_apply", gout.to(DEV), out, yd, mean, invstd, gamma.to(DEV), None, None, sums, dy,
         dgamma, dbeta, rows, C, 0, 0)
    dg2, db2 = dgamma.clone(), dbeta.clone()
    call("b2n_bn_bwd_apply", gout.to(DEV), out, yd, mean, invstd, gamma.to(DEV), None, None, sums, dy,
         dg2, db2, rows, C, 0, 1)                               # accumulate into a gradient slot
    assert torch.equal(dg2, 2 * dgamma) and torch.equal(db2, 2 * dbeta)
    scale_t = float(yr.grad.abs().max())
    assert float((dy.cpu() - yr.grad).abs().max()) < 2e-4 * scale_t
    assert torch.allclose(dgamma.cpu(), bn.weight.grad, rtol=1e-3, atol=1e-3)
    assert torch.allclose(dbeta.cpu(), bn.bias.grad, rtol=1e-4, atol=1e-4)
    # ReLU gate recomputed from y (no residual: bn1 of a block) == gate read from the mask tensor
    act = torch.empty(rows, C, device=DEV)
    call("b2n_bn_apply", yd, scale, shift, None, None, None, None, None, act, None, None, rows, C, 1, 0)
    s_m, s_y = (torch.zeros(2 * C, device=DEV, dtype=torch.float64) for _ in range(2))
    call("b2n_bn_bwd_reduce", gout.to(DEV), act, yd, mean, invstd, None, None, s_m, rows, C)
    call("b2n_bn_bwd_reduce", gout.to(DEV), None, yd, mean, invstd, scale, shift, s_y, rows, C)
    assert torch.equal(s_m, s_y)
    dy_m, dy_y = torch.empty(rows, C, device=DEV), torch.empty(rows, C, device=DEV)
    call("b2n_bn_bwd_apply", gout.to(DEV), act, yd, mean, invstd, gamma.to(DEV), None, None, s_m, dy_m,
         dgamma, dbeta, rows, C, 1, 0)
    call("b2n_bn_bwd_apply", gout.to(DEV), None, yd, mean, invstd, gamma.to(DEV), scale, shift, s_m, dy_y,
         dgamma, dbeta, rows, C, 1, 0)
    assert torch.equal(dy_m, dy_y)
    # eval-mode fold
    call("b2n_bn_fold_eval", gamma.to(DEV), beta.to(DEV), rm, rv, scale, shift, C, 1e-5)
    bn.eval()
    call("b2n_bn_apply", yd, scale, shift, None, None, None, None, None, out, None, None, rows, C, 0,
         0)
    assert torch.allclose(out.cpu(), bn(y).detach(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 7, 9, 64), (3, 112, 112, 64)])
def test_bn_relu_maxpool_forward_backward(shape):
    N, H, W, C = shape
    g = torch.Generator().manual_seed(H * W)
    y = torch.randn(N, C, H, W, generator=g)
    scale, shift = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    yr = y.clone().requires_grad_(True)
    z = torch.relu(yr * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    a_ref = F.max_pool2d(z, 3, 2, 1)
    z.retain_grad()
    ga = torch.randn(a_ref.shape, generator=g)
    a_ref.backward(ga)
    P, Q = a_ref.shape[2:]
    a = torch.empty(N, P, Q, C, device=DEV)
    idx = torch.empty(N, P, Q, C, device=DEV, dtype=torch.uint8)
    yd = to_nhwc(y).to(DEV)
    a_h, a_l = torch.empty_like(a, dtype=torch.half), torch.empty_like(a, dtype=torch.half)
    call("b2n_bn_relu_maxpool", yd, scale.to(DEV), shift.to(DEV), a, a_h, a_l, idx, N, H, W, C)
    assert torch.allclose(from_nhwc((a_h.float() + a_l.float()).cpu()), a_ref.detach(), rtol=1e-5,
                          atol=1e-6)
    assert torch.allclose(from_nhwc(a.cpu()), a_ref.detach(), rtol=1e-3, atol=1e-6)
    assert torch.equal(a.cpu(), O.tf32_round(a.cpu()))
    gz = torch.empty(N, H, W, C, device=DEV)
    call("b2n_maxpool_relu_bwd", to_nhwc(ga).to(DEV), idx, yd, scale.to(DEV), shift.to(DEV), gz, N, H,
         W, C)
    assert torch.allclose(from_nhwc(gz.cpu()), z.grad * (z.detach() > 0), rtol=1e-6, atol=1e-7)
    # fused maxpool + ReLU + BN backward == the three-kernel chain over the materialised gz
    rows = N * H * W
    gamma = (torch.rand(C, generator=g) + 0.5).to(DEV)
    mean = yd.mean((0, 1, 2))
    invstd = 1.0 / torch.sqrt(yd.var((0, 1, 2), unbiased=False) + 1e-5)
    s_ref = torch.zeros(2 * C, device=DEV, dtype=torch.float64)
    call("b2n_bn_bwd_reduce", gz, None, yd, mean, invstd, None, None, s_ref, rows, C)
    dy_ref, dg_ref, db_ref = torch.empty_like(yd), torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    call("b2n_bn_bwd_apply", gz, None, yd, mean, invstd, gamma, None, None, s_ref, dy_ref, dg_ref, db_ref,
         rows, C, 1, 0)
    s_f = torch.zeros(2 * C, device=DEV, dtype=torch.float64)
    call("b2n_pool_bn_bwd_reduce", to_nhwc(ga).to(DEV), idx, yd, scale.to(DEV), shift.to(DEV), mean,
         invstd, s_f, N, H, W, C)
    assert torch.allclose(s_f, s_ref, rtol=1e-5, atol=1e-5)
    dy_f, dg_f, db_f = torch.empty_like(yd), torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    call("b2n_pool_bn_bwd_apply", to_nhwc(ga).to(DEV), idx, yd, scale.to(DEV), shift.to(DEV), mean,
         invstd, gamma, s_ref, dy_f, dg_f, db_f, N, H, W, C, 1, 0)
    # (a position picked by 3-4 windows sums their gradients in scatter order: a last-bit
    # difference there can move the TF32 rounding of dy by one TF32 ulp = 2^-11 relative)
    assert torch.allclose(dy_f, dy_ref, rtol=1e-3, atol=1e-6 * float(dy_ref.abs().max()))
    assert float((dy_f != dy_ref).float().mean()) < 1e-3
    assert torch.equal(dg_f, dg_ref) and torch.equal(db_f, db_ref)


def test_multi_tensor_packs_and_folds_equal_the_single_launches():
    """b2n_pack_weights_multi / b2n_bn_fold_eval_multi (one launch for all conv weights / BatchNorm
    layers of a pass) write exactly what the per-tensor entry points write -- also past the 64-job
    (32-layer) tables of one launch."""
    import ctypes
    g = torch.Generator().manual_seed(5)
    shapes = [(64, 64, 3, 3), (128, 64, 3, 3), (128, 64, 1, 1), (256, 128, 3, 3), (64, 32, 3, 3)] * 14   # 70 jobs
    ws, kinds, outs, refs = [], [], [], []
    for i, (K, C, R, S) in enumerate(shapes):
        w = (torch.randn(K, C, R, S, generator=g) * 0.05).to(DEV)
        kind = i % 3 if R == 3 else i % 2
        ws.append(w)
        kinds.append(kind)
        if kind == 0:
            o, r = (torch.empty(2, K, R * S * C, device=DEV, dtype=torch.float16) for _ in range(2))
            call("b2n_pack_weight_fwd", w, r[0], r[1], K, C, R, S)
            outs.append((o[0], o[1]))
        elif kind == 1:
            o, r = (torch.empty(C, R * S * K, device=DEV) for _ in range(2))
            call("b2n_pack_weight_dgrad", w, r, K, C, R, S)
            outs.append((o, o))
        else:
            o, r = (torch.empty(C, 9 * K, device=DEV) for _ in range(2))
            call("b2n_pack_weight_dgrad_s2m", w, r, K, C)
            outs.append((o, o))
        refs.append(r)
    n = len(ws)
    PtrArr, IntArr = ctypes.c_void_p * n, ctypes.c_int * n
    call("b2n_pack_weights_multi", PtrArr(*[w.data_ptr() for w in ws]), PtrArr(*[o[0].data_ptr() for o in outs]),
         PtrArr(*[o[1].data_ptr() for o in outs]), IntArr(*kinds), IntArr(*[w.shape[0] for w in ws]),
         IntArr(*[w.shape[1] for w in ws]), IntArr(*[w.shape[2] for w in ws]), IntArr(*[w.shape[3] for w in ws]),
         n, device=torch.device(DEV))
    for kind, o, r in zip(kinds, outs, refs):
        if kind == 0:
            assert torch.equal(o[0], r[0]) and torch.equal(o[1], r[1])
        else:
            assert torch.equal(o[0], r)
    # BatchNorm folds: 40 layers of mixed widths
    Cs = [64, 128, 256, 512, 96] * 8
    n = len(Cs)
    par = [[(torch.rand(C, generator=g) + 0.5).to(DEV) for C in Cs] for _ in range(4)]   # gamma, beta, rm, rv
    sc, sh = [torch.empty(C, device=DEV) for C in Cs], [torch.empty(C, device=DEV) for C in Cs]
    PtrArr, IntArr, FltArr = ctypes.c_void_p * n, ctypes.c_int * n, ctypes.c_float * n
    eps = [1e-5 if i % 2 else 1e-3 for i in range(n)]
    call("b2n_bn_fold_eval_multi", *[PtrArr(*[t.data_ptr() for t in ts]) for ts in par],
         PtrArr(*[t.data_ptr() for t in sc]), PtrArr(*[t.data_ptr() for t in sh]), IntArr(*Cs), FltArr(*eps), n,
         device=torch.device(DEV))
    for i, C in enumerate(Cs):
        rs, rh = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
        call("b2n_bn_fold_eval", par[0][i], par[1][i], par[2][i], par[3][i], rs, rh, C, eps[i])
        assert torch.equal(sc[i], rs) and torch.equal(sh[i], rh)


def test_avgpool():
    a = torch.randn(5, 49, 512)
    e = torch.empty(5, 512, device=DEV)
    ah, al = split_pair(a)
    call("b2n_avgpool_fwd", ah.to(DEV), al.to(DEV), e, 5, 49, 512)
    assert torch.allclose(e.cpu(), a.mean(1), rtol=1e-5, atol=1e-6)
    ge = torch.randn(5, 512)
    g = torch.empty(5, 49, 512, device=DEV)
    call("b2n_avgpool_bwd", ge.to(DEV), None, g, 5, 49, 512)
    want = (ge / 49).unsqueeze(1).expand(5, 49, 512)
    assert torch.allclose(g.cpu(), want, rtol=1e-6)
    call("b2n_avgpool_bwd", ge.to(DEV), a.to(DEV), g, 5, 49, 512)      # gated by the pooled activation
    assert torch.allclose(g.cpu(), torch.where(a > 0, want, torch.zeros(())), rtol=1e-6)


@pytest.mark.parametrize("n", [1, 2, 77, 768])
def test_heads_forward_backward_fp32(n):
    from ssl_cr_histo_b200 import heads

    torch.manual_seed(n)
    l1, l2 = torch.nn.Linear(1024, 512), torch.nn.Linear(512, 256)
    x = torch.randn(n, 1024)
    dy = torch.randn(n, 256)
    xr = x.clone().requires_grad_(True)
    ref = l2(torch.relu(l1(xr)))
    ref.backward(dy)
    m1, m2 = torch.nn.Linear(1024, 512).to(DEV), torch.nn.Linear(512, 256).to(DEV)
    m1.load_state_dict(l1.state_dict()); m2.load_state_dict(l2.state_dict())
    xm = x.to(DEV).requires_grad_(True)
    out = heads.mlp2(xm, m1, m2)
    out.backward(dy.to(DEV))
    assert torch.allclose(out.cpu(), ref.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(xm.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-5)
    for a, b in ((m1, l1), (m2, l2)):
        assert torch.allclose(a.weight.grad.cpu(), b.weight.grad, rtol=1e-4, atol=1e-4)
        assert torch.allclose(a.bias.grad.cpu(), b.bias.grad, rtol=1e-4, atol=1e-4)
    lin, lin_m = torch.nn.Linear(768, 9), torch.nn.Linear(768, 9).to(DEV)
    lin_m.load_state_dict(lin.state_dict())
    f = torch.randn(n, 768)
    assert torch.allclose(heads.linear(f.to(DEV), lin_m).cpu(), lin(f).detach(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("n", [1, 5, 256])
def test_pair_mlp_matches_the_concatenating_reference(n):
    """models/net.py:56-64 (three concatenated pairs through the shared MLP, concatenated again) and
    its E1 == E2 == E3 special case (:88-103), forward and every gradient -- also when the parameter
    gradients are delivered into arena slots that already hold a value (accumulate)."""
    from ssl_cr_histo_b200 import heads

    torch.manual_seed(100 + n)
    fc = torch.nn.Sequential(torch.nn.Linear(1024, 512), torch.nn.ReLU(True), torch.nn.Linear(512, 256))
    E = [torch.randn(n, 512, requires_grad=True) for _ in range(3)]
    dy = torch.randn(n, 768)
    ref = O._pairwise_features(fc, *E)
    ref.backward(dy)
    m = torch.nn.Sequential(torch.nn.Linear(1024, 512), torch.nn.ReLU(True), torch.nn.Linear(512, 256)).to(DEV)
    m.load_state_dict(fc.state_dict())
    Em = [e.detach().to(DEV).requires_grad_(True) for e in E]
    out = heads.pair_mlp(*Em, m[0], m[2])
    out.backward(dy.to(DEV))
    assert torch.allclose(out.cpu(), ref.detach(), rtol=1e-4, atol=1e-5)
    for a, b in zip(Em, E):
        assert torch.allclose(a.grad.cpu(), b.grad, rtol=1e-4, atol=1e-5)
    for a, b in zip(m.parameters(), fc.parameters()):
        assert torch.allclose(a.grad.cpu(), b.grad, rtol=1e-4, atol=1e-4)
    # the same input three times, evaluated once; gradients into pre-filled slots
    for p in fc.parameters():
        p.grad = None
    e = torch.randn(n, 512, requires_grad=True)
    ref = O._pairwise_features(fc, e, e, e)
    ref.backward(dy)
    for p in m.parameters():
        p.grad = None
        p._b2n_grad_slot = torch.ones_like(p)
    em = e.detach().to(DEV).requires_grad_(True)
    out = heads.pair_mlp_same(em, m[0], m[2])
    out.backward(dy.to(DEV))
    assert torch.allclose(out.cpu(), ref.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(em.grad.cpu(), e.grad, rtol=1e-4, atol=1e-5)
    for a, b in zip(m.parameters(), fc.parameters()):
        assert a.grad is None                                   # autograd saw nothing: it went to the slot
        assert torch.allclose(a._b2n_grad_slot.cpu() - 1, b.grad, rtol=1e-4, atol=1e-4)


def test_fused_losses_match_torch():
    g = torch.Generator().manual_seed(9)
    # RSP cross-entropy + argmax
    lg = torch.randn(37, 6, generator=g)
    tgt = torch.randint(0, 6, (37,), generator=g)
    lr = lg.clone().requires_grad_(True)
    ref = F.cross_entropy(lr, tgt)
    ref.backward()
    lm = lg.to(DEV).requires_grad_(True)
    loss, pred = losses.cross_entropy(lm, tgt.to(DEV))
    loss.backward()
    assert abs(float(loss) - float(ref)) < 1e-6 * max(1, abs(float(ref)))
    assert torch.equal(pred.cpu(), torch.argmax(lg, 1))
    assert torch.allclose(lm.grad.cpu(), lr.grad, rtol=1e-5, atol=1e-7)
    # BreastPathQ MSE / MSE
    lx, tx = torch.randn(12, 1, generator=g), torch.rand(12, generator=g)
    lw, ls = torch.randn(20, 1, generator=g), torch.randn(20, 1, generator=g)
    a, b = lx.clone().requires_grad_(True), ls.clone().requires_grad_(True)
    ref = F.mse_loss(a, tx.view(-1, 1)) + 0.7 * F.mse_loss(lw, b)
    ref.backward()
    am, bm = lx.to(DEV).requires_grad_(True), ls.to(DEV).requires_grad_(True)
    total, parts = losses.consistency_mse(am, tx.to(DEV), lw.to(DEV), bm, 0.7)
    total.backward()
    assert abs(float(total) - float(ref)) < 1e-6 and abs(float(parts[2]) - float(ref)) < 1e-6
    assert torch.allclose(am.grad.cpu(), a.grad, rtol=1e-5, atol=1e-8)
    assert torch.allclose(bm.grad.cpu(), b.grad, rtol=1e-5, atol=1e-8)
    # Kather CE + CE-to-teacher-argmax
    lx, tx = torch.randn(9, 9, generator=g), torch.randint(0, 9, (9,), generator=g)
    lw, ls = torch.randn(24, 9, generator=g), torch.randn(24, 9, generator=g)
    a, b = lx.clone().requires_grad_(True), ls.clone().requires_grad_(True)
    tu = torch.max(torch.softmax(lw, -1), -1)[1]
    ref = F.cross_entropy(a, tx) + 1.0 * F.cross_entropy(b, tu)
    ref.backward()
    am, bm = lx.to(DEV).requires_grad_(True), ls.to(DEV).requires_grad_(True)
    total, parts, pred, pseudo = losses.consistency_ce(am, tx.to(DEV), lw.to(DEV), bm, 1.0)
    total.backward()
    assert abs(float(total) - float(ref)) < 1e-6 * max(1, abs(float(ref)))
    assert torch.equal(pseudo.cpu(), tu) and torch.equal(pred.cpu(), torch.argmax(lx, 1))
    assert torch.allclose(am.grad.cpu(), a.grad, rtol=1e-5, atol=1e-7)
    assert torch.allclose(bm.grad.cpu(), b.grad, rtol=1e-5, atol=1e-7)
    # a label outside [0, C) is never dereferenced: the loss turns NaN instead of reading garbage
    bad = tgt.clone(); bad[5] = -100
    loss, _ = losses.cross_entropy(lg.to(DEV), bad.to(DEV))
    assert torch.isnan(loss)
    with pytest.raises(RuntimeError, match="must both be"):
        losses.consistency_ce(lx.to(DEV), tx.to(DEV), lw[:20].to(DEV), ls.to(DEV), 1.0)
    with pytest.raises(RuntimeError, match="class targets"):
        losses.cross_entropy(lg.to(DEV), tgt[:30].to(DEV))


def test_lerp_handoff_bitwise_and_lookahead_golden():
    from util import golden

    g = golden("lookahead.npz")
    shapes = [(7,), (3, 5), (2, 2, 2)]
    flat, grads = torch.tensor(g["init"]), torch.tensor(g["grads"])
    params, gs, off = [], [], 0
    for s in shapes:
        n = int(np.prod(s))
        params.append(flat[off:off + n].clone().view(s).to(DEV)); gs.append(grads[off:off + n].view(s).to(DEV))
        off += n
    cached = [p.clone() for p in params]
    for step in range(7):
        for p, gr in zip(params, gs):
            p.add_(gr, alpha=-0.1)
        if (step + 1) % 5 == 0:
            weights.lookahead_pull_(params, cached, 0.5)
            for p, c in zip(params, cached):
                assert torch.equal(p, c)
        now = torch.cat([p.flatten() for p in params]).cpu()
        ref = torch.tensor(g["trace"][step])
        assert float((now - ref).abs().max()) <= 1.2e-7 * float(ref.abs().max()), step  # <= 1 ulp
    # alpha = 1 is a bit-exact copy even over inf / nan destinations
    dst = [torch.full((1000,), float("nan"), device=DEV), torch.full((3, 3), float("inf"), device=DEV)]
    src = [torch.randn(1000, device=DEV), torch.randn(3, 3, device=DEV)]
    weights.lerp_(dst, src, 1.0)
    assert all(torch.equal(d, s) for d, s in zip(dst, src))
    # empty list and > 96 tensors (table chunking)
    weights.lerp_([], [], 0.5)
    many_d = [torch.zeros(5, device=DEV) for _ in range(130)]
    many_s = [torch.full((5,), float(i), device=DEV) for i in range(130)]
    weights.lerp_(many_d, many_s, 0.25)
    assert all(torch.allclose(d, s * 0.25) for d, s in zip(many_d, many_s))


@pytest.mark.parametrize("kind", ["adam", "sgd_nesterov", "sgd_plain"])
def test_fused_optimizers_match_torch_optim(kind):
    """b2n_adam_multi / b2n_sgd_multi follow torch.optim's formulas term by term: three steps over a
    ragged list of 70 tensors (> one 64-tensor launch), a frozen parameter, weight decay, an LR
    change in between, and grad_scale == pre-scaled gradients."""
    from ssl_cr_histo_b200 import optim as fused
    g = torch.Generator().manual_seed(11)
    shapes = [(64, 3, 7, 7), (64,), (128, 64, 3, 3), (1,), (512, 1024)] + [(5 + i, 3) for i in range(65)]
    init = [torch.randn(s, generator=g) for s in shapes]
    mine = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    ref = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    if kind == "adam":
        om = fused.Adam(mine, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
        orf = torch.optim.Adam(ref, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    else:
        kw = dict(lr=0.05, momentum=0.9 if kind == "sgd_nesterov" else 0.0, weight_decay=1e-4,
                  nesterov=kind == "sgd_nesterov")
        om, orf = fused.SGD(mine, **kw), torch.optim.SGD(ref, **kw)
    om.grad_scale = 0.25
    for step in range(3):
        for i, (a, b) in enumerate(zip(mine, ref)):
            if i == 3:
                continue                                     # a parameter that never gets a grad
            gr = torch.randn(a.shape, generator=g).to(DEV)
            a.grad, b.grad = gr * 4.0, gr.clone()            # grad_scale 0.25 undoes the x4 exactly
        om.step(); orf.step()
        if step == 0:
            om.param_groups[0]["lr"] *= 0.1; orf.param_groups[0]["lr"] *= 0.1
    for i, (a, b) in enumerate(zip(mine, ref)):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), (kind, i, float((a - b).abs().max()))
    assert torch.equal(mine[3], ref[3])
    if kind == "adam":
        assert om.state[mine[0]]["step"] == 3
        assert torch.allclose(om.state[mine[0]]["exp_avg_sq"], orf.state[ref[0]]["exp_avg_sq"], rtol=2e-6)


def test_capturable_adam_counts_steps_on_the_device_and_replays_from_a_graph():
    """optim.Adam(capturable=True): the bias corrections come from a device-side step counter that
    the launch itself advances -- three eager steps and three CUDA-graph replays of one captured
    step() equal six torch.optim.Adam steps."""
    from ssl_cr_histo_b200 import optim as fused
    g = torch.Generator().manual_seed(13)
    shapes = [(64, 3, 7, 7), (64,), (33, 5)] + [(7 + i,) for i in range(70)]
    init = [torch.randn(s, generator=g) for s in shapes]
    mine = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    ref = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    om = fused.Adam(mine, lr=1e-2, weight_decay=1e-4, capturable=True)
    orf = torch.optim.Adam(ref, lr=1e-2, weight_decay=1e-4)
    grads = [torch.randn(s, generator=g).to(DEV) for s in shapes]
    for a, b, gr in zip(mine, ref, grads):
        a.grad, b.grad = gr.clone(), gr.clone()                # static gradient buffers
    for _ in range(3):
        om.step(); orf.step()
    for i, (a, b) in enumerate(zip(mine, ref)):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), i
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        om.step()
    for k in range(3):
        for a, b in zip(mine, ref):                            # new gradient values, same buffers
            a.grad.mul_(0.5 + k); b.grad.mul_(0.5 + k)
        graph.replay(); orf.step()
    for i, (a, b) in enumerate(zip(mine, ref)):
        assert torch.allclose(a, b, rtol=4e-6, atol=2e-7), (i, float((a - b).abs().max()))


def test_fused_optimizer_invalidates_packed_weights():
    """A step writes parameters behind autograd's version counters: the trunk must re-pack."""
    import ssl_cr_histo_b200.net as net
    from ssl_cr_histo_b200 import optim as fused
    torch.manual_seed(0)
    m = net.TripletNet_Finetune("resnet18").to(DEV).train()
    x = ints((2, 3, 32, 32), 0, 255, 5).to(DEV)
    opt = fused.SGD(m.parameters(), lr=0.1, momentum=0.9, nesterov=True)
    out0 = m(x)
    out0.square().mean().backward()
    opt.step()
    with torch.no_grad():
        out1 = m(x)
    assert not torch.equal(out0, out1)


@pytest.mark.parametrize("case", [(2, 56, 56, 64, 64, 3, 1, 1), (3, 13, 9, 64, 64, 3, 1, 1)])
def test_tap_sharing_kernels_equal_the_per_tap_kernels(case, monkeypatch):
    """The HALO variants (one TMA box per filter row, taps through row-shifted descriptors) and the
    generic one-box-per-tap kernels compute the same sums: forward (both operand modes) and weight
    gradient, bit for bit on exactly representable data (B2N_NO_HALO selects the generic path)."""
    N, H, W, Cin, Cout, R, s, p = case
    x = ints((N, Cin, H, W), -4, 4, 21)
    w = ints((Cout, Cin, R, R), -2, 2, 22, 0.25)
    dy = ints((N, Cout, H, W), -2, 2, 23, 0.5)
    xh, xl = split_pair(to_nhwc(x))

    def run():
        y1 = conv(to_nhwc(x).to(DEV), pack_fwd(w), N, H, W, Cin, Cout, R, s, p, p)
        y2 = conv((xh.to(DEV), xl.to(DEV)), pack_fwd(w, split=True), N, H, W, Cin, Cout, R, s, p, p)
        dwp = torch.zeros(Cout, R * R * Cin, device=DEV)
        call("b2n_conv_wgrad", to_nhwc(x).to(DEV), to_nhwc(dy).to(DEV), dwp, N, H, W, Cin, Cout, R, R, s,
             p, p, p, p, 0, 0)
        return y1, y2, dwp

    halo = run()
    monkeypatch.setenv("B2N_NO_HALO", "1")
    plain = run()
    ref = F.conv2d(x, w, None, s, p)
    assert torch.equal(from_nhwc(halo[0].cpu()), ref)
    for a, b in zip(halo, plain):
        assert torch.equal(a, b)
