"""ctypes binding of libb2n.so (the C ABI declared in include/b2n.h).

There is no CPU fallback: if the library cannot be loaded, or an entry point reports an
error, a RuntimeError is raised.  Tensors are passed as raw device pointers; every call is
enqueued on ``torch.cuda.current_stream()``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2n.so")

P, I, LL, F, D = c_void_p, c_int, c_longlong, c_float, c_double

# name -> argument ctypes (the trailing stream argument is appended automatically)
_SIGNATURES = {
    "b2n_conv_fwd": [P] * 9 + [I] * 12 + [P, P, P, P, P, P, I, I, P, P] + [I] * 5 + [P] * 6,
    "b2n_conv_wgrad": [P, P, P] + [I] * 14,
    "b2n_pack_weight_fwd": [P, P, P, I, I, I, I],
    "b2n_pack_weight_dgrad": [P, P, I, I, I, I],
    "b2n_pack_weights_multi": [P] * 8 + [I],
    "b2n_pack_weight_dgrad_s2": [P, P, I, I],
    "b2n_pack_weight_dgrad_s2m": [P, P, I, I],
    "b2n_conv_dgrad_s2": [P, P, P, I, I, I, I, I, I, I, P, P],
    "b2n_conv_dgrad_s2_sc": [P, P, P, P, P, I, I, I, I, I, I, I, P],
    "b2n_unpack_wgrad": [P, P, I, I, I, I, I, I],
    "b2n_stem_pack_input": [P, P, P, P, P, I, I, I],
    "b2n_stem_pack_input_u8": [P, P, P, I, I, I],
    "b2n_stem_pack_weight": [P, P, P, I],
    "b2n_stem_unpack_wgrad": [P, P, I, I, I],
    "b2n_bn_finalize": [P] * 10 + [I, D, F, F, I],
    "b2n_bn_fold_eval": [P] * 6 + [I, F],
    "b2n_bn_fold_eval_multi": [P] * 8 + [I],
    "b2n_bn_apply": [P] * 11 + [LL, I, I, I],
    "b2n_bn_bwd_reduce": [P] * 8 + [LL, I],
    "b2n_bn_bwd_apply": [P] * 12 + [LL, I, I, I],
    "b2n_upsample_zero": [P, P] + [I] * 6,
    "b2n_bn_relu_maxpool": [P] * 7 + [I] * 4,
    "b2n_maxpool_relu_bwd": [P] * 6 + [I] * 4,
    "b2n_pool_bn_bwd_reduce": [P] * 8 + [I] * 4,
    "b2n_pool_bn_bwd_apply": [P] * 12 + [I] * 6,
    "b2n_avgpool_fwd": [P, P, P, I, I, I],
    "b2n_avgpool_bwd": [P, P, P, I, I, I],
    "b2n_linear_fwd": [P, LL, P, LL, P, P, LL, I, I, I, I, I],
    "b2n_linear_bwd_data": [P, LL, P, LL, P, LL, P, I, I, I, I],
    "b2n_linear_bwd_weight": [P, LL, P, LL, P, LL, P, I, I, I, I],
    "b2n_cols_replicate": [P, LL, I, I, I],
    "b2n_cols_sum": [P, LL, P, I, I, I],
    "b2n_fused_loss": [I, P, P, P, P, P, I, I, I, F, P, P, P, P, P],
    "b2n_softmax_last": [P, P, I, I],
    "b2n_lerp_multi": [P, P, P, I, F, I],
    "b2n_adam_multi": [P, P, P, P, P, I, D, D, D, D, D, LL, P, D],
    "b2n_sgd_multi": [P, P, P, P, I, D, D, D, I, I, D],
    "b2n_aug_flip_crop": [P, P, P, P, P, I, I, I, I, I],
    "b2n_aug_brightness_contrast": [P, P, P, P, P, I, I, I],
    "b2n_aug_image_mean": [P, P, I, I, I],
    "b2n_aug_hsv_shift": [P, P, P, P, P, P, I, I, I],
    "b2n_aug_add_noise": [P, P, P, P, I, I, I],
    "b2n_aug_box_blur": [P, P, P, P, I, I, I],
    "b2n_aug_hed_jitter": [P, P, P, P, I, I, I],
    "b2n_aug_warp_affine": [P, P, P, P, I, I, I, I, I, I],
}
EXPORTS = ["b2n_version", "b2n_last_error", "b2n_device_ok", "b2n_launch_count",
           "b2n_conv_wgrad_planes"] + list(_SIGNATURES)

_lib = None
# bumped whenever a kernel writes parameters behind autograd's back (lerp), so that cached
# weight packs keyed on tensor._version are invalidated
WEIGHT_EPOCH = 0


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError("libb2n.so is missing (%s); run __graft_entry__.build()" % LIB_PATH)
        from . import build as _build

        _build.build_lib()      # serialised across processes by a file lock (one rank compiles)
    lib = ctypes.CDLL(LIB_PATH)
    lib.b2n_version.restype = c_int
    lib.b2n_last_error.restype = c_char_p
    lib.b2n_device_ok.restype = c_int
    lib.b2n_launch_count.restype = ctypes.c_ulonglong
    lib.b2n_conv_wgrad_planes.restype = c_int
    lib.b2n_conv_wgrad_planes.argtypes = [c_int] * 12
    for name, sig in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = list(sig) + [c_void_p]
        fn.restype = c_int
    _lib = lib
    return lib


def _conv(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a


# Optional per-kernel timing (bench.py): when PROFILE is a dict, calls whose entry-point name
# is a key get bracketed by CUDA events on the launching stream; `work` (algorithmic FLOPs or
# bytes of that launch -- or a tuple (algorithmic FLOPs, executed FP16-MMA FLOPs, executed
# TF32-MMA FLOPs, tag) -- supplied by the caller) is recorded next to them.
PROFILE = None


def call(name: str, *args, work=0.0, device=None) -> None:
    """Invoke an entry point on the current CUDA stream of the device that owns the tensor
    arguments (``device``: for entry points that take pointer tables instead of tensors) --
    switching device for the call if it is not the current one, since the library's launches and
    per-device function attributes follow cudaGetDevice; raise on a non-zero return code."""
    if device is None:
        for a in args:
            if isinstance(a, torch.Tensor):
                device = a.device
                break
    if device is not None and device.type == "cuda" and device.index is not None and \
            device.index != torch.cuda.current_device():
        with torch.cuda.device(device):
            return _call(name, args, work)
    return _call(name, args, work)


def _call(name: str, args, work) -> None:
    lib = load()
    stream = torch.cuda.current_stream().cuda_stream
    prof = PROFILE.get(name) if PROFILE is not None else None
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*[_conv(a) for a in args], stream)
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (name, lib.b2n_last_error().decode()))
    if prof is not None:
        e1.record()
        prof.append((e0, e1, work))


def wgrad_planes(*shape) -> int:
    """Partial planes a deterministic b2n_conv_wgrad of this shape writes (host-side query)."""
    lib = load()
    n = lib.b2n_conv_wgrad_planes(*[int(v) for v in shape])
    if n < 1:
        raise RuntimeError("b2n_conv_wgrad_planes failed: %s" % lib.b2n_last_error().decode())
    return n


def launch_count() -> int:
    return int(load().b2n_launch_count())


def require_device(t: torch.Tensor, what: str = "input") -> None:
    if not t.is_cuda:
        raise RuntimeError(
            "ssl_cr_histo_b200 runs only on CUDA (sm_100a) tensors; got a %s %s -- there is no "
            "CPU fallback" % (t.device.type, what))
