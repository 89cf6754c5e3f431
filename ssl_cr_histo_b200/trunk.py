"""ResNet18 trunk (fc removed) executed by the libb2n CUDA kernels.

Drop-in for the ``torchvision.models.resnet18(pretrained=False)`` + ``model.fc = Sequential()``
object that models/net.py:32-34 builds: same parameter / buffer names, order, shapes and
initialisation, same train / eval BatchNorm semantics, autograd-visible.
``nn.Conv2d`` / ``nn.BatchNorm2d`` objects are used only as parameter containers -- their
``forward`` is never called; all arithmetic goes through ``_lib.call``.

Layout: activations NHWC fp32 in HBM.  Per conv+BN unit the raw conv output ``y`` (FP32
accumulators), and the normalised / activated tensor ``a`` (TF32-rounded, because it is the
next conv's tensor-core operand) are kept for the backward pass.
"""
from __future__ import annotations

import os
import threading
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import call

_LAYER_CFG = [(64, 1), (128, 2), (256, 2), (512, 2)]  # (planes, stride of first block)
STEM_C = 32    # reduction channels per tap of the stem's weight gradient (12 real)
STEM_C_STORED = 12  # channels the fp32 space-to-depth copy stores; the TMA unit zero-fills the rest
STEM_C16 = 16  # channels of its FP16 (hi, lo) pair, the forward conv's operand


class BasicBlock(nn.Module):
    """Parameter container with torchvision's attribute names (conv1, bn1, conv2, bn2,
    downsample.0/.1).  Registration order follows torchvision's BasicBlock.__init__ so that
    ``named_parameters()`` enumerates identically (index-based freezing relies on it)."""

    def __init__(self, inplanes: int, planes: int, stride: int, downsample: Optional[nn.Module]):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class _PackCache:
    """Packed (K-major) copies of conv weights.  Never copied or saved.

    An entry is rebuilt when its tag -- the parameter's storage, autograd version counter,
    ``_b2n_epoch`` (bumped by this package's optimizer kernels) or the global ``WEIGHT_EPOCH``
    (bumped by ``weights.lerp_``) -- changes.  Writes made through ``p.data`` (the vendored
    Lookahead, lookahead.py:96-97; RAdam; any EMA update) leave all of those untouched, so in
    addition every pack is rebuilt on the first training-mode forward after a backward pass (the
    point where an optimizer of any kind has had its turn) and after ``invalidate()``.  An
    eval-mode module whose weights are written through ``.data`` outside this package must call
    ``trunk.invalidate_packs()``."""

    def __init__(self):
        self.entries = {}
        self.floor = 0          # entries built before this generation are stale
        self.generation = 1

    def __deepcopy__(self, memo):
        return _PackCache()

    def __getstate__(self):        # torch.save(model): the packs are derived data, never pickled
        return {}

    def __setstate__(self, state):
        self.__init__()

    def invalidate(self):
        self.generation += 1
        self.floor = self.generation

    @staticmethod
    def _tag(param: torch.Tensor):
        return (param.data_ptr(), param._version, getattr(param, "_b2n_epoch", 0), _lib.WEIGHT_EPOCH,
                param.device)

    def stale(self, key: str, param: torch.Tensor) -> bool:
        hit = self.entries.get(key)
        return hit is None or hit[0] != self._tag(param) or hit[2] < self.floor

    def put(self, key: str, param: torch.Tensor, packed) -> None:
        self.entries[key] = (self._tag(param), packed, self.generation)

    def get(self, key: str, param: torch.Tensor, maker):
        if not self.stale(key, param):
            return self.entries[key][1]
        packed = maker(param.detach())
        self.put(key, param, packed)
        return packed


def _pack_fwd(w):
    """(hi, lo) FP16 pair of the forward operand [K][(r*S+s)*C + c]."""
    K, C, R, S = w.shape
    out = torch.empty(2, K, R * S * C, device=w.device, dtype=torch.float16)
    call("b2n_pack_weight_fwd", w, out[0], out[1], K, C, R, S)
    return out[0], out[1]


def _pack_dgrad(w):
    K, C, R, S = w.shape
    out = torch.empty(C, R * S * K, device=w.device, dtype=torch.float32)
    call("b2n_pack_weight_dgrad", w, out, K, C, R, S)
    return out


def _pack_dgrad_s2(w):
    """Four parity-class packs of a stride-2 3x3 data gradient (views of one buffer)."""
    K, C = w.shape[0], w.shape[1]
    buf = torch.empty(9 * C * K, device=w.device, dtype=torch.float32)
    call("b2n_pack_weight_dgrad_s2", w, buf, K, C)
    out, off = [], 0
    for ntap in (1, 2, 2, 4):
        out.append(buf[off:off + C * ntap * K].view(C, ntap * K))
        off += C * ntap * K
    return out


def _pack_dgrad_s2m(w):
    """The nine (class, tap) blocks of a stride-2 3x3 data gradient as one [C][9*K] matrix: the
    weight operand of the merged launch (b2n_conv_dgrad_s2)."""
    K, C = w.shape[0], w.shape[1]
    out = torch.empty(C, 9 * K, device=w.device, dtype=torch.float32)
    call("b2n_pack_weight_dgrad_s2m", w, out, K, C)
    return out


def _pack_stem(w):
    out = torch.empty(2, w.shape[0], 16 * STEM_C16, device=w.device, dtype=torch.float16)
    call("b2n_stem_pack_weight", w, out[0], out[1], w.shape[0])
    return out[0], out[1]


def _prefill_packs(trunk: "ResNet18Trunk", with_dgrad: bool) -> None:
    """Build every stale weight pack of the eight blocks in ONE launch (b2n_pack_weights_multi)
    instead of one 2-8 us kernel per pack as the pass reaches it: a training step repacks all
    twenty conv weights after its optimizer step (~40 packs with the data-gradient layouts), a
    graph-captured pass always does.  Same keys / layouts as the lazy ``packs.get`` calls below,
    which then hit.  (The stem's space-to-depth pack stays a launch of its own.)"""
    import ctypes

    packs = trunk._packs
    jobs = []   # (key, weight, kind, value to cache, dst0, dst1)

    def add(key, conv, kind):
        w = conv.weight
        if not packs.stale(key, w):
            return
        K, C, R, S = w.shape
        if kind == 0:
            out = torch.empty(2, K, R * S * C, device=w.device, dtype=torch.float16)
            jobs.append((key, w, 0, (out[0], out[1]), out[0], out[1]))
        elif kind == 1:
            out = torch.empty(C, R * S * K, device=w.device, dtype=torch.float32)
            jobs.append((key, w, 1, out, out, out))
        else:
            out = torch.empty(C, 9 * K, device=w.device, dtype=torch.float32)
            jobs.append((key, w, 2, out, out, out))

    for bi, blk in enumerate(trunk.blocks()):
        add("b%d.w1" % bi, blk.conv1, 0)
        add("b%d.w2" % bi, blk.conv2, 0)
        if blk.downsample is not None:
            add("b%d.wd" % bi, blk.downsample[0], 0)
        if with_dgrad:
            add("b%d.w2d" % bi, blk.conv2, 1)
            if blk.downsample is not None:
                add("b%d.wdd" % bi, blk.downsample[0], 1)
                if MERGED_S2_DGRAD:
                    add("b%d.w1s2m" % bi, blk.conv1, 2)
            else:
                add("b%d.w1d" % bi, blk.conv1, 1)
    n = len(jobs)
    if n == 0:
        return
    PtrArr, IntArr = ctypes.c_void_p * n, ctypes.c_int * n
    call("b2n_pack_weights_multi",
         PtrArr(*[j[1].data_ptr() for j in jobs]), PtrArr(*[j[4].data_ptr() for j in jobs]),
         PtrArr(*[j[5].data_ptr() for j in jobs]), IntArr(*[j[2] for j in jobs]),
         IntArr(*[j[1].shape[0] for j in jobs]), IntArr(*[j[1].shape[1] for j in jobs]),
         IntArr(*[j[1].shape[2] for j in jobs]), IntArr(*[j[1].shape[3] for j in jobs]), n,
         device=jobs[0][1].device)
    for key, w, _kind, value, _d0, _d1 in jobs:
        packs.put(key, w, value)


def _fold_eval_all(bns, bufs) -> None:
    """Eval mode: fold the running statistics of every BatchNorm layer into its (scale, shift)
    slot of ``bufs`` in one launch (b2n_bn_fold_eval_multi)."""
    import ctypes

    n = len(bns)
    PtrArr, IntArr, FltArr = ctypes.c_void_p * n, ctypes.c_int * n, ctypes.c_float * n
    offs, o = [], 0
    for b in bns:
        offs.append(o)
        o += b.num_features
    es = bufs.element_size()
    call("b2n_bn_fold_eval_multi",
         PtrArr(*[b.weight.data_ptr() for b in bns]), PtrArr(*[b.bias.data_ptr() for b in bns]),
         PtrArr(*[b.running_mean.data_ptr() for b in bns]), PtrArr(*[b.running_var.data_ptr() for b in bns]),
         PtrArr(*[bufs[0].data_ptr() + es * x for x in offs]), PtrArr(*[bufs[1].data_ptr() + es * x for x in offs]),
         IntArr(*[b.num_features for b in bns]), FltArr(*[b.eps for b in bns]), n, device=bufs.device)


class ResNet18Trunk(nn.Module):
    """(N,3,H,W) NCHW in [0,255], fp32 (as the reference's loops feed it) or uint8 ->
    (N,512) features.  H and W must be even."""

    def __init__(self):
        super().__init__()
        # Construction order and RNG consumption mirror torchvision.models.resnet.ResNet.__init__
        # (resnet.py:197-213) so that the same torch seed yields the same initial weights.
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        for li, (planes, stride) in enumerate(_LAYER_CFG, start=1):
            blocks = []
            for bi in range(2):
                s = stride if bi == 0 else 1
                down = None
                if s != 1 or inplanes != planes:
                    down = nn.Sequential(nn.Conv2d(inplanes, planes, 1, s, bias=False),
                                         nn.BatchNorm2d(planes))
                blocks.append(BasicBlock(inplanes, planes, s, down))
                inplanes = planes
            setattr(self, "layer%d" % li, nn.Sequential(*blocks))
        nn.Linear(512, 1000)  # torchvision builds (and the reference then discards) this fc
        self.fc = nn.Sequential()
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self._packs = _PackCache()

    # ------------------------------------------------------------------ plumbing
    def blocks(self) -> List[BasicBlock]:
        return [b for li in (1, 2, 3, 4) for b in getattr(self, "layer%d" % li)]

    def bn_layers(self) -> List[nn.BatchNorm2d]:
        out = [self.bn1]
        for b in self.blocks():
            out += [b.bn1, b.bn2]
            if b.downsample is not None:
                out.append(b.downsample[1])
        return out

    def invalidate_packs(self) -> None:
        """Drop the cached weight packs.  Only needed after writing parameters of an *eval-mode*
        module through ``p.data`` (which no version counter sees); training-mode forwards refresh
        the packs after every backward pass on their own."""
        self._packs.invalidate()

    def forward(self, x: torch.Tensor, n_updates: int = 1) -> torch.Tensor:
        """``n_updates`` > 1 applies that many identical BN running-stat updates in closed form
        (TripletNet_Finetune runs the same input through the trunk three times)."""
        _lib.require_device(x)
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError("expected (N,3,H,W) input, got %s" % (tuple(x.shape),))
        params = [p for p in self.parameters()]
        return _TrunkFn.apply(x, self, n_updates, torch.is_grad_enabled(), *params)


# ====================================================================== execution


# Weight gradients are off the backward pass' critical path (nothing downstream reads them until
# the optimizer), so they run on a side stream next to the HBM-bound kernels of the main chain.
# B2N_OVERLAP_WGRAD=2 (default) is the ordered schedule: a weight gradient is held back until the
# next element-wise phase of the main chain (BatchNorm-backward reduce / apply, the pooling
# sweeps), launched on the side stream just before it, and the next tensor-core launch of the main
# chain waits for it -- a weight gradient only ever shares the SMs with HBM-bound kernels, never
# with another persistent conv kernel.  Measured on the power-capped B200s: -0.5 ms per cfg3 step
# at one GPU, -0.4 ms at two (gradients bit-identical to the single-stream pass, also through the
# overlapped bucket all-reduce: tools/ddp_check.py).  =1 is the first, unordered variant (the weight
# gradient starts as soon as its operands exist and may share the SMs with the next data-gradient
# conv: -0.2 ms on average, occasional +3..+20 ms outliers when two persistent 1-CTA/SM kernels
# contend), =0 keeps the whole pass on one stream.
_SIDE_STREAMS = {}
# Per-thread state of the backward pass that is running (nn.DataParallel drives one host thread per
# GPU, so two passes may be in flight in one process): `before_conv` -- hooks called before every
# tensor-core launch of the main chain; `streams` -- the (main, side) streams of the pass when it
# uses a side stream: a consumer that is started from inside the pass (ddp.GradAllReducer's bucket
# all-reduce) must wait for both.
_TLS = threading.local()


def _before_conv_hooks() -> list:
    h = getattr(_TLS, "before_conv", None)
    if h is None:
        h = _TLS.before_conv = []
    return h


def active_backward_streams() -> list:
    st = getattr(_TLS, "streams", None)
    if st is None:
        st = _TLS.streams = []
    return st


_ov = os.environ.get("B2N_OVERLAP_WGRAD", "2")
OVERLAP_WGRAD = 0 if _ov in ("", "0") else (2 if _ov == "2" else 1)
# Stride-2 data gradients: one merged launch over dY for the four output-parity classes (default),
# or -- B2N_NO_S2M=1 -- one launch per class (same arithmetic; dY is then read four times).
MERGED_S2_DGRAD = os.environ.get("B2N_NO_S2M", "0") in ("", "0")
# ... with the block's 1x1 shortcut gradient accumulated inside that launch (default), or --
# B2N_NO_S2M_SC=1 -- as a 1x1 launch of its own whose result the merged kernel adds (resid).
FUSED_S2_SHORTCUT = os.environ.get("B2N_NO_S2M_SC", "0") in ("", "0")

# Weight gradients are split-K sums over pixel slabs.  By default every split stores its own plane
# and the unpack kernel adds the planes in a fixed order: bit-repeatable gradients for ~20 MB of
# transient scratch per conv and 0.1 ms of a 34 ms step.  B2N_DETERMINISTIC=0 (or
# set_deterministic(False)) reduces the partial tiles with fp32 atomics into one zeroed plane
# instead (summation order, hence the last bits, then vary from run to run).
DETERMINISTIC_WGRAD = os.environ.get("B2N_DETERMINISTIC", "1") not in ("", "0")


def set_deterministic(flag: bool) -> None:
    """Bit-repeatable weight gradients (two-stage split-K reduction instead of atomics)."""
    global DETERMINISTIC_WGRAD
    DETERMINISTIC_WGRAD = bool(flag)


def _side_stream(dev: torch.device) -> torch.cuda.Stream:
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return s


class _BNState:
    __slots__ = ("scale", "shift", "mean", "invstd", "inv_gamma")


def _bn_affine(bn: nn.BatchNorm2d, training: bool, stats, count, n_updates, bufs, slot,
               want_inv_gamma=False):
    """Per-channel affine for one BN layer; in training mode also the running-stat update."""
    C = bn.num_features
    st = _BNState()
    st.scale, st.shift, st.mean, st.invstd = (bufs[i][slot:slot + C] for i in range(4))
    st.inv_gamma = torch.empty(C, device=bufs.device, dtype=torch.float32) if want_inv_gamma else None
    if training:
        momentum = 0.1 if bn.momentum is None else bn.momentum
        call("b2n_bn_finalize", stats, bn.weight, bn.bias, bn.running_mean, bn.running_var,
             st.scale, st.shift, st.mean, st.invstd, st.inv_gamma, C, float(count), momentum, bn.eps,
             n_updates)
    # (eval mode: the slots were filled for all layers at once, _fold_eval_all)
    return st


class _Act:
    """A forward activation: the error-compensated (hi, lo) FP16 pair the next forward conv
    multiplies (hi = fp16(a), lo = fp16(a - hi)) and, only when the backward pass will need it,
    the TF32-rounded fp32 copy ``f32`` (wgrad operand and ReLU mask)."""
    __slots__ = ("hi", "lo", "f32")

    def __init__(self, shape, dev, with_f32):
        self.hi = torch.empty(shape, device=dev, dtype=torch.float16)
        self.lo = torch.empty(shape, device=dev, dtype=torch.float16)
        self.f32 = torch.empty(shape, device=dev, dtype=torch.float32) if with_f32 else None


def _conv(x, wp, N, H, W, Cin, Cout, R, stride, pad_lo, pad_hi, *, scale=None, shift=None,
          resid=None, resid_pair=None, mask=None, relu=0, rnd=0, stats=None, out=None,
          out_pair=None, want_out=True, alg=1.0, lo_flag=None, pad_hi_w=None, R_w=None,
          placement=(0, 0, 0, 0, 0), gate=None, bnb=None):
    """One conv launch.  ``x`` / ``wp`` are either an ``_Act`` and an FP16 (hi, lo) weight pair
    (error-compensated forward) or plain fp32 tensors (single TF32 pass: data gradients).
    ``alg``: algorithmic / executed FLOP ratio of this launch (the stem runs 147 real taps in a
    256-wide padded reduction) -- only used for the roofline accounting in bench.py, which gets
    (algorithmic FLOPs, executed FP16-MMA FLOPs, executed TF32-MMA FLOPs, tag) per launch."""
    S = R if R_w is None else R_w                 # taps / upper padding may differ per axis
    phw = pad_hi if pad_hi_w is None else pad_hi_w
    P = (H + pad_lo + pad_hi - R) // stride + 1
    Q = (W + pad_lo + phw - S) // stride + 1
    if isinstance(x, _Act):
        x32, xh, xl, w32, (wh, wl), dev = None, x.hi, x.lo, None, wp, x.hi.device
    else:
        x32, xh, xl, w32, wh, wl, dev = x, None, None, wp, None, None, x.device
    if out is None and want_out:
        out = torch.empty(N, P, Q, Cout, device=dev, dtype=torch.float32)
    oh, ol = (out_pair.hi, out_pair.lo) if out_pair is not None else (None, None)
    rh, rl = (resid_pair.hi, resid_pair.lo) if resid_pair is not None else (None, None)
    nominal = 2.0 * N * P * Q * Cout * R * S * Cin
    # algorithmic DRAM bytes of the launch: every operand / result tensor touched once
    n_out = float(N) * P * Q * Cout if placement[0] == 0 else float(N) * P * Q * Cout
    byt = n_out * (4.0 * (out is not None) + 4.0 * (out_pair is not None) + 4.0 * (resid is not None)
                   + 4.0 * (mask is not None) + 4.0 * (resid_pair is not None)
                   + 4.0 * (gate is not None) + 4.0 * (bnb is not None))
    if isinstance(x, _Act):   # hi*hi + hi*lo + lo*hi (the lo plane of integer images is skipped)
        byt += float(N) * H * W * Cin * (2.0 if lo_flag is not None else 4.0)
        work = (nominal * alg, nominal * (2.0 if lo_flag is not None else 3.0), 0.0, "fwd", byt)
    else:
        byt += float(N) * H * W * Cin * 4.0
        work = (nominal * alg, 0.0, nominal, "dgrad", byt)
    # bnb = (y, BN state, gate_from_y): BatchNorm-backward sums of the result, see b2n.h
    by, bst, bgate = bnb if bnb is not None else (None, None, False)
    hooks = _before_conv_hooks()       # (backward pass with ordered side-stream weight gradients)
    if hooks:
        hooks[-1]()
    call("b2n_conv_fwd", x32, xh, xl, w32, wh, wl, out, oh, ol, N, H, W, Cin, Cout, R, S, stride,
         pad_lo, pad_hi, pad_lo, phw, scale, shift, resid, rh, rl, mask, relu, rnd, stats,
         lo_flag, *placement, gate, by, bst.mean if bst else None, bst.invstd if bst else None,
         bst.scale if bgate else None, bst.shift if bgate else None, work=work)
    return out


class _TrunkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, trunk: ResNet18Trunk, n_updates: int, grad_mode: bool, *params):
        training = trunk.training
        # grad mode is always off inside Function.forward and needs_input_grad ignores
        # torch.no_grad(), so the caller's grad mode is passed in explicitly
        any_grad = grad_mode and any(ctx.needs_input_grad[4:])
        if any_grad and not training:
            raise NotImplementedError(
                "eval-mode BatchNorm with trainable trunk parameters is not on the reference's "
                "path (teacher / validation run under no_grad); call .train() or no_grad()")
        save = any_grad
        dev = x.device
        # uint8 pixels (the patches before the loop's .float(), dataset.py:65-67) are accepted
        # as they are: 4x fewer host->device bytes, and exact in FP16 (no lo plane)
        u8 = x.dtype == torch.uint8
        x = x.contiguous() if u8 else x.contiguous().float()
        N, _, H, W = x.shape
        if (H | W) & 1:
            raise RuntimeError("H and W must be even (got %dx%d)" % (H, W))
        packs = trunk._packs
        if (training and getattr(trunk, "_packs_dirty", True)) or \
                torch.cuda.is_current_stream_capturing():
            # see _PackCache: p.data writes are invisible to the tags.  Under CUDA-graph capture the
            # pack kernels must be part of the graph, whatever the cache holds.
            packs.invalidate()
            trunk._packs_dirty = False
        _prefill_packs(trunk, with_dgrad=save and all(ctx.needs_input_grad[4:]))
        bns = trunk.bn_layers()
        total_c = sum(b.num_features for b in bns)
        stats_all = torch.zeros(2 * total_c, device=dev, dtype=torch.float64) if training else None
        bufs = torch.empty(4, total_c, device=dev, dtype=torch.float32)
        if not training:
            _fold_eval_all(bns, bufs)    # slot order = bn_layers() order = the next_bn() calls below
        slot = [0]

        def next_bn(bn):
            C = bn.num_features
            s = slot[0]
            slot[0] += C
            stats = stats_all[2 * s:2 * s + 2 * C] if training else None
            return s, stats

        saved = {"N": N, "H": H, "W": W, "blocks": []}
        stem_alg = 147.0 / (16 * STEM_C16)

        # ---- stem: s2d pack -> 4x4 tensor-core conv -> BN+ReLU+maxpool
        H2, W2 = H // 2, W // 2
        xs = _Act.__new__(_Act)
        xs.hi = torch.empty(N, H2, W2, STEM_C16, device=dev, dtype=torch.float16)
        xs.f32 = torch.empty(N, H2, W2, STEM_C_STORED, device=dev, dtype=torch.float32) if save else None
        lo_flag = torch.zeros(1, device=dev, dtype=torch.int32)
        if u8:
            xs.lo = xs.hi           # never read: lo_flag stays 0
            call("b2n_stem_pack_input_u8", x, xs.hi, xs.f32, N, H, W)
        else:
            xs.lo = torch.empty_like(xs.hi)
            call("b2n_stem_pack_input", x, xs.hi, xs.lo, xs.f32, lo_flag, N, H, W)
        ws = packs.get("stem", trunk.conv1.weight, _pack_stem)
        s0, stats0 = next_bn(trunk.bn1)
        PH, PW = (H2 - 1) // 2 + 1, (W2 - 1) // 2 + 1
        a = _Act((N, PH, PW, 64), dev, save)
        if training:
            y0 = _conv(xs, ws, N, H2, W2, STEM_C16, 64, 4, 1, 2, 1, stats=stats0, alg=stem_alg,
                       lo_flag=lo_flag)
            bn0 = _bn_affine(trunk.bn1, True, stats0, N * H2 * W2, n_updates, bufs, s0,
                             want_inv_gamma=save)
            idx = torch.empty(N, PH, PW, 64, device=dev, dtype=torch.uint8) if save else None
            call("b2n_bn_relu_maxpool", y0, bn0.scale, bn0.shift, a.f32, a.hi, a.lo, idx, N, H2, W2,
                 64)
            if save:
                saved.update(xs=xs.f32, y0=y0, bn0=bn0, idx=idx)
        else:
            bn0 = _bn_affine(trunk.bn1, False, None, 0, 0, bufs, s0)
            z0 = _conv(xs, ws, N, H2, W2, STEM_C16, 64, 4, 1, 2, 1, scale=bn0.scale, shift=bn0.shift,
                       relu=1, alg=stem_alg, lo_flag=lo_flag)
            ones = torch.ones(64, device=dev)
            zeros = torch.zeros(64, device=dev)
            call("b2n_bn_relu_maxpool", z0, ones, zeros, None, a.hi, a.lo, None, N, H2, W2, 64)
        h, w = PH, PW

        # ---- eight basic blocks
        for bi, blk in enumerate(trunk.blocks()):
            cin, cout, s = blk.conv1.in_channels, blk.conv1.out_channels, blk.stride
            ph, pw = (h + 2 - 3) // s + 1, (w + 2 - 3) // s + 1
            w1 = packs.get("b%d.w1" % bi, blk.conv1.weight, _pack_fwd)
            w2 = packs.get("b%d.w2" % bi, blk.conv2.weight, _pack_fwd)
            rows = N * ph * pw
            rec = {"a_in": a.f32, "h": h, "w": w, "ph": ph, "pw": pw}
            s1, st1 = next_bn(blk.bn1)
            s2, st2 = next_bn(blk.bn2)
            if blk.downsample is not None:
                wd = packs.get("b%d.wd" % bi, blk.downsample[0].weight, _pack_fwd)
                sd, std = next_bn(blk.downsample[1])
            a1 = _Act((N, ph, pw, cout), dev, save)
            a_out = _Act((N, ph, pw, cout), dev, save)
            if training:
                y1 = _conv(a, w1, N, h, w, cin, cout, 3, s, 1, 1, stats=st1)
                b1 = _bn_affine(blk.bn1, True, st1, rows, n_updates, bufs, s1)
                call("b2n_bn_apply", y1, b1.scale, b1.shift, None, None, None, None, None, a1.f32,
                     a1.hi, a1.lo, rows, cout, 1, 1)
                y2 = _conv(a1, w2, N, ph, pw, cout, cout, 3, 1, 1, 1, stats=st2)
                b2 = _bn_affine(blk.bn2, True, st2, rows, n_updates, bufs, s2)
                if blk.downsample is not None:
                    yd = _conv(a, wd, N, h, w, cin, cout, 1, s, 0, 0, stats=std)
                    bd = _bn_affine(blk.downsample[1], True, std, rows, n_updates, bufs, sd)
                    call("b2n_bn_apply", y2, b2.scale, b2.shift, yd, bd.scale, bd.shift, None, None,
                         a_out.f32, a_out.hi, a_out.lo, rows, cout, 1, 1)
                    rec.update(yd=yd, bd=bd)
                else:
                    call("b2n_bn_apply", y2, b2.scale, b2.shift, None, None, None, a.hi, a.lo,
                         a_out.f32, a_out.hi, a_out.lo, rows, cout, 1, 1)
                rec.update(y1=y1, a1=a1.f32, y2=y2, a_out=a_out.f32, b1=b1, b2=b2)
            else:
                # eval: BN folded into the conv epilogue, no intermediate tensors
                b1 = _bn_affine(blk.bn1, False, None, 0, 0, bufs, s1)
                b2 = _bn_affine(blk.bn2, False, None, 0, 0, bufs, s2)
                _conv(a, w1, N, h, w, cin, cout, 3, s, 1, 1, scale=b1.scale, shift=b1.shift, relu=1,
                      out_pair=a1, want_out=False)
                if blk.downsample is not None:
                    bd = _bn_affine(blk.downsample[1], False, None, 0, 0, bufs, sd)
                    idn = _conv(a, wd, N, h, w, cin, cout, 1, s, 0, 0, scale=bd.scale,
                                shift=bd.shift)
                    idn_pair = None
                else:
                    idn, idn_pair = None, a
                _conv(a1, w2, N, ph, pw, cout, cout, 3, 1, 1, 1, scale=b2.scale, shift=b2.shift,
                      resid=idn, resid_pair=idn_pair, relu=1, out_pair=a_out, want_out=False)
            if save:
                saved["blocks"].append(rec)
            a, h, w = a_out, ph, pw

        e = torch.empty(N, 512, device=dev, dtype=torch.float32)
        call("b2n_avgpool_fwd", a.hi, a.lo, e, N, h * w, 512)
        if training:    # nn.BatchNorm2d's per-forward counter, all twenty layers in one launch
            torch._foreach_add_([b.num_batches_tracked for b in bns], n_updates)
        ctx.trunk = trunk
        ctx.saved = saved if save else None
        ctx.bufs = bufs
        return e

    @staticmethod
    def backward(ctx, ge):
        trunk, sv = ctx.trunk, ctx.saved
        if sv is None:
            raise RuntimeError("trunk backward called without saved activations")
        ctx.saved = None  # free activations as early as possible
        trunk._packs_dirty = True   # an optimizer follows: the next training forward repacks
        params = list(trunk.parameters())
        needs = ctx.needs_input_grad[4:]
        grads = {id(p): None for p in params}
        need = {id(p): n for p, n in zip(params, needs)}
        dev = ge.device
        N = sv["N"]
        packs = trunk._packs
        blocks = trunk.blocks()

        def unit_needs(mods):
            return any(need[id(p)] for m in mods for p in m.parameters())

        unit_need = [unit_needs([trunk.conv1, trunk.bn1])] + [unit_needs([b]) for b in blocks]
        if not any(unit_need):
            return (None, None, None, None) + tuple(None for _ in params)
        min_unit = min(i for i, n in enumerate(unit_need) if n)

        # One zero-fill per backward pass for every BatchNorm-backward sum pair and every packed
        # weight-gradient accumulator (instead of one torch.zeros launch each).
        bn_list = trunk.bn_layers()
        sums_all = torch.zeros(2 * sum(b.num_features for b in bn_list), device=dev, dtype=torch.float64)
        sums_off = {}
        off = 0
        for b in bn_list:
            sums_off[id(b)] = off
            off += 2 * b.num_features

        def sums_of(bn):
            o = sums_off[id(bn)]
            return sums_all[o:o + 2 * bn.num_features]

        convs = [trunk.conv1] + [c for b in blocks for c in
                                 ([b.conv1, b.conv2] + ([b.downsample[0]] if b.downsample is not None else []))]
        det = DETERMINISTIC_WGRAD
        dwp_off, total = {}, 0
        for c in convs:
            if need[id(c.weight)] and not det:
                dwp_off[id(c)] = total
                # the stem's packed gradient is [64][16 taps * STEM_C] (space-to-depth view)
                total += 64 * 16 * STEM_C if c is trunk.conv1 else c.weight.numel()
        dwp_all = torch.zeros(total, device=dev, dtype=torch.float32) if total else None

        def dwp_of(conv, rows, cols):
            o = dwp_off[id(conv)]
            return dwp_all[o:o + rows * cols].view(rows, cols)

        def emit(p, make):
            """Deliver a parameter gradient: into the parameter's arena slot (ddp.GradAllReducer,
            accumulating -- autograd then sees None and launches nothing), else as a fresh tensor."""
            slot = getattr(p, "_b2n_grad_slot", None)
            if slot is not None:
                make(slot, 1)
                sink = getattr(p, "_b2n_grad_sink", None)
                if sink is not None:
                    sink.ready(p)
            else:
                t = torch.empty_like(p)
                make(t, 0)
                grads[id(p)] = t

        main = torch.cuda.current_stream(dev)
        # (per-launch timing for bench.py's roofline keeps everything on one stream)
        side = _side_stream(dev) if OVERLAP_WGRAD and _lib.PROFILE is None else None
        # Ordered overlap (B2N_OVERLAP_WGRAD=2): weight gradients wait in `pending` until the next
        # element-wise phase, run on the side stream next to it, and the next tensor-core launch of
        # the main chain waits for them (side_busy = their completion event).
        ordered = side is not None and OVERLAP_WGRAD == 2
        pending, side_busy = [], [None]

        def elementwise_phase():
            """Called right before the main chain launches HBM-bound kernels: the held-back weight
            gradients go to the side stream first (their single CTA per SM is placed before the
            element-wise blocks fill the thread slots)."""
            if not pending:
                return
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                for fn in pending:
                    fn()
                done = torch.cuda.Event()
                done.record(side)
            side_busy[0] = done
            pending.clear()

        def before_conv():
            if side_busy[0] is not None:
                main.wait_event(side_busy[0])
                side_busy[0] = None

        hooks, active = _before_conv_hooks(), active_backward_streams()
        del hooks[:]              # (a hook left behind by a pass that raised is dropped here)
        del active[:]
        if ordered:
            hooks.append(before_conv)
        if side is not None:
            active.extend((main, side))

        def bn_backward(g, y, st, bn, rows, C, reduced=False, gate_from_y=False):
            """g: gradient w.r.t. the BN output, already ReLU-gated by its producer (the gate of
            every activation gradient is applied by the kernel that writes it).  ``reduced``: the
            producer also accumulated the two BatchNorm-backward sums (conv epilogue).
            ``gate_from_y``: g is not gated yet and the ReLU input is this BN's own output (the
            stem's unfused chain)."""
            gsc, gsh = (st.scale, st.shift) if gate_from_y else (None, None)
            sums = sums_of(bn)
            if ordered:
                elementwise_phase()
            if not reduced:
                call("b2n_bn_bwd_reduce", g, None, y, st.mean, st.invstd, gsc, gsh, sums, rows, C)
            dy = torch.empty_like(y)
            wg, bg = need[id(bn.weight)], need[id(bn.bias)]
            sw = getattr(bn.weight, "_b2n_grad_slot", None) if wg else None
            sb = getattr(bn.bias, "_b2n_grad_slot", None) if bg else None
            if wg and bg and sw is not None and sb is not None:
                dgamma, dbeta, acc = sw, sb, 1
            else:
                dgamma, dbeta, acc = torch.empty(C, device=dev), torch.empty(C, device=dev), 0
            call("b2n_bn_bwd_apply", g, None, y, st.mean, st.invstd, bn.weight, gsc, gsh, sums, dy,
                 dgamma, dbeta, rows, C, 1, acc)
            if acc:
                for p in (bn.weight, bn.bias):
                    sink = getattr(p, "_b2n_grad_sink", None)
                    if sink is not None:
                        sink.ready(p)
            else:
                if wg:
                    grads[id(bn.weight)] = dgamma
                if bg:
                    grads[id(bn.bias)] = dbeta
            return dy

        class _on_side:
            """Run the enclosed launches on the side stream once everything enqueued on the main
            stream so far (the operands) is complete; ``uses``: main-stream tensors they read."""

            def __init__(self, *uses):
                self.uses = uses

            def __enter__(self):
                if side is None:
                    return
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                for t in self.uses:
                    t.record_stream(side)
                self.ctx = torch.cuda.stream(side)
                self.ctx.__enter__()

            def __exit__(self, *exc):
                if side is not None:
                    self.ctx.__exit__(*exc)

        def wgrad(conv, x_in, dy, H, W, stride, pad):
            if not need[id(conv.weight)]:
                return
            K, C, R, S = conv.weight.shape
            P, Q = dy.shape[1], dy.shape[2]
            wflops = 2.0 * N * P * Q * K * R * S * C

            def run():
                planes = _lib.wgrad_planes(N, H, W, C, K, R, S, stride, pad, pad, pad, pad) if det else 1
                dwp = torch.empty(planes, K, R * S * C, device=dev) if det else dwp_of(conv, K, R * S * C)
                call("b2n_conv_wgrad", x_in, dy, dwp, N, H, W, C, K, R, S, stride, pad, pad, pad, pad,
                     1 if det else 0, 0,
                     work=(wflops, 0.0, wflops, "wgrad", 4.0 * N * (H * W * C + P * Q * K)))
                emit(conv.weight, lambda t, acc: call("b2n_unpack_wgrad", dwp, t, K, C, R, S, acc, planes))
                dw = grads[id(conv.weight)]
                if side is not None and dw is not None:
                    dw.record_stream(main)   # consumed by autograd / the optimizer on the main stream

            if ordered:
                def deferred():
                    x_in.record_stream(side)
                    dy.record_stream(side)
                    run()
                pending.append(deferred)
                return
            with _on_side(x_in, dy):
                run()

        # The stem's maxpool + ReLU + BN backward runs as two band sweeps over the stem output (below)
        # when two double-buffered band slots (2 rows of y + 2 pooled rows of gradients / codes) fit
        # in shared memory: patches up to ~300 px wide; wider ones take the unfused chain.  Its
        # reduction sweep is not needed at all: xhat at a window's argmax is (a - beta) / gamma of
        # the pooled activation a, so the first block's data-gradient epilogue takes both sums.
        H2s, W2s = sv["H"] // 2, sv["W"] // 2
        q2 = (W2s - 1) // 2 + 1
        stem_fused = min_unit == 0 and 2 * (2 * W2s * 64 * 4 + 2 * q2 * 64 * 5) + 1024 <= 227 * 1024
        stem_sums_done = False

        last = sv["blocks"][-1]
        hw = last["ph"] * last["pw"]
        g = torch.empty(N, last["ph"], last["pw"], 512, device=dev, dtype=torch.float32)
        # g is the gradient w.r.t. the block's post-ReLU output: gated here, at its producer
        call("b2n_avgpool_bwd", ge.contiguous().float(), last["a_out"], g, N, hw, 512)

        for bi in range(len(blocks) - 1, -1, -1):
            unit = bi + 1
            if unit < min_unit:
                break
            blk, rec = blocks[bi], sv["blocks"][bi]
            cin, cout, s = blk.conv1.in_channels, blk.conv1.out_channels, blk.stride
            h, w, ph, pw = rec["h"], rec["w"], rec["ph"], rec["pw"]
            rows = N * ph * pw
            need_in = unit > min_unit
            # the block input is a post-ReLU block output (gate its gradient) except for the first
            # block, whose input is the max-pooled stem activation (the pool backward gates)
            in_gate = rec["a_in"] if bi > 0 else None
            # main branch: bn2 <- conv2 <- relu/bn1 <- conv1
            dy2 = bn_backward(g, rec["y2"], rec["b2"], blk.bn2, rows, cout)
            wgrad(blk.conv2, rec["a1"], dy2, ph, pw, 1, 1)
            wd2 = packs.get("b%d.w2d" % bi, blk.conv2.weight, _pack_dgrad)
            # conv2's data gradient: bn1's ReLU gate (recomputed from y1) and both BatchNorm-backward
            # sums are taken in the conv epilogue -- no separate reduction pass over da1 / y1
            da1 = _conv(dy2, wd2, N, ph, pw, cout, cout, 3, 1, 1, 1, stats=sums_of(blk.bn1),
                        bnb=(rec["y1"], rec["b1"], True))
            dy1 = bn_backward(da1, rec["y1"], rec["b1"], blk.bn1, rows, cout, reduced=True)
            wgrad(blk.conv1, rec["a_in"], dy1, h, w, s, 1)
            g_in = None
            if blk.downsample is not None:
                dconv, dbn = blk.downsample[0], blk.downsample[1]
                dyd = bn_backward(g, rec["yd"], rec["bd"], dbn, rows, cout)
                wgrad(dconv, rec["a_in"], dyd, h, w, s, 0)
                if need_in:
                    # stride-2 data gradients by output parity (no zero-stuffing): the 1x1
                    # shortcut conv only reaches even pixels; the 3x3 conv is four small stride-1
                    # tap subsets over dY, each writing (and ReLU-gating) its own quarter of g_in.
                    g_in = torch.empty(N, h, w, cin, device=dev, dtype=torch.float32)
                    wdd = packs.get("b%d.wdd" % bi, dconv.weight, _pack_dgrad)
                    fused_sc = MERGED_S2_DGRAD and FUSED_S2_SHORTCUT
                    if not fused_sc:
                        _conv(dyd, wdd, N, ph, pw, cout, cin, 1, 1, 0, 0, out=g_in,
                              placement=(2, 0, 0, h, w))
                    if MERGED_S2_DGRAD:
                        wm = packs.get("b%d.w1s2m" % bi, blk.conv1.weight, _pack_dgrad_s2m)
                        flops = 2.0 * N * ph * pw * cout * cin * (10 if fused_sc else 9)
                        byt = 4.0 * N * (ph * pw * cout * (2 if fused_sc else 1)
                                         + h * w * cin * (2 if in_gate is not None else 1)
                                         + (0 if fused_sc else ph * pw * cin))
                        before_conv()
                        if fused_sc:
                            call("b2n_conv_dgrad_s2_sc", dy1, wm, dyd, wdd, g_in, N, ph, pw, cout, cin, h, w,
                                 in_gate, work=(flops, 0.0, flops, "dgrad", byt))
                        else:
                            call("b2n_conv_dgrad_s2", dy1, wm, g_in, N, ph, pw, cout, cin, h, w, g_in, in_gate,
                                 work=(flops, 0.0, flops, "dgrad", byt))
                    else:
                        wcls = packs.get("b%d.w1s2" % bi, blk.conv1.weight, _pack_dgrad_s2)
                        for cls, (a0, b0) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                            _conv(dy1, wcls[cls], N, ph, pw, cout, cin, 1 + a0, 1, 0, a0, R_w=1 + b0,
                                  pad_hi_w=b0, out=g_in, resid=g_in if cls == 0 else None,
                                  placement=(2, a0, b0, h, w), gate=in_gate)
            elif need_in:
                wd1 = packs.get("b%d.w1d" % bi, blk.conv1.weight, _pack_dgrad)
                # identity shortcut: add the (already gated) upstream gradient in the epilogue
                if bi == 0 and stem_fused and sv["bn0"].inv_gamma is not None:
                    pooled = _BNState()
                    pooled.mean, pooled.invstd = trunk.bn1.bias, sv["bn0"].inv_gamma
                    g_in = _conv(dy1, wd1, N, ph, pw, cout, cin, 3, 1, 1, 1, resid=g, gate=rec["a_in"],
                                 stats=sums_of(trunk.bn1), bnb=(rec["a_in"], pooled, False))
                    stem_sums_done = True
                else:
                    g_in = _conv(dy1, wd1, N, ph, pw, cout, cin, 3, 1, 1, 1, resid=g, gate=in_gate)
            g = g_in

        if min_unit == 0:
            H2, W2 = H2s, W2s
            bn0 = sv["bn0"]
            if stem_fused:
                # maxpool + ReLU + BN backward fused: one sweep over the stem output (two when the
                # sums did not come out of the first block's data-gradient epilogue) instead of five
                sums = sums_of(trunk.bn1)
                if ordered:
                    elementwise_phase()
                if not stem_sums_done:
                    call("b2n_pool_bn_bwd_reduce", g, sv["idx"], sv["y0"], bn0.scale, bn0.shift,
                         bn0.mean, bn0.invstd, sums, N, H2, W2, 64)
                dy0 = torch.empty_like(sv["y0"])
                bw, bb = trunk.bn1.weight, trunk.bn1.bias
                sw = getattr(bw, "_b2n_grad_slot", None) if need[id(bw)] else None
                sb = getattr(bb, "_b2n_grad_slot", None) if need[id(bb)] else None
                if sw is not None and sb is not None:
                    dgamma, dbeta, acc = sw, sb, 1
                else:
                    dgamma, dbeta, acc = torch.empty(64, device=dev), torch.empty(64, device=dev), 0
                call("b2n_pool_bn_bwd_apply", g, sv["idx"], sv["y0"], bn0.scale, bn0.shift, bn0.mean,
                     bn0.invstd, bw, sums, dy0, dgamma, dbeta, N, H2, W2, 64, 1, acc)
                if acc:
                    for p in (bw, bb):
                        sink = getattr(p, "_b2n_grad_sink", None)
                        if sink is not None:
                            sink.ready(p)
                else:
                    if need[id(bw)]:
                        grads[id(bw)] = dgamma
                    if need[id(bb)]:
                        grads[id(bb)] = dbeta
            else:
                gz = torch.empty_like(sv["y0"])
                if ordered:
                    elementwise_phase()
                call("b2n_maxpool_relu_bwd", g, sv["idx"], sv["y0"], bn0.scale, bn0.shift, gz, N, H2,
                     W2, 64)
                dy0 = bn_backward(gz, sv["y0"], bn0, trunk.bn1, N * H2 * W2, 64)
            if need[id(trunk.conv1.weight)]:
                with _on_side(sv["xs"], dy0):
                    planes = _lib.wgrad_planes(N, H2, W2, STEM_C, 64, 4, 4, 1, 2, 1, 2, 1) if det else 1
                    dws = (torch.empty(planes, 64, 16 * STEM_C, device=dev) if det
                           else dwp_of(trunk.conv1, 64, 16 * STEM_C))
                    call("b2n_conv_wgrad", sv["xs"], dy0, dws, N, H2, W2, STEM_C, 64, 4, 4, 1, 2, 1, 2, 1,
                         1 if det else 0, STEM_C_STORED,
                         work=(2.0 * N * H2 * W2 * 64 * 147, 0.0, 2.0 * N * H2 * W2 * 64 * 16 * STEM_C,
                               "wgrad", 4.0 * N * H2 * W2 * (STEM_C_STORED + 64)))
                    emit(trunk.conv1.weight,
                         lambda t, acc: call("b2n_stem_unpack_wgrad", dws, t, 64, acc, planes))
                dw = grads[id(trunk.conv1.weight)]
                if side is not None and dw is not None:
                    dw.record_stream(main)

        if ordered:
            elementwise_phase()      # weight gradients still held back (nothing left to pair them with)
            del hooks[:]
        if side is not None:
            main.wait_stream(side)   # all weight gradients are complete before anyone reads them
            del active[:]
        return (None, None, None, None) + tuple(grads[id(p)] for p in params)
