"""Whole-model weight interpolation in one kernel launch (csrc/loss_lerp.cu, b2n_lerp_multi).

    dst <- alpha * src + (1 - alpha) * dst        over every (dst, src) tensor pair

* ``teacher_handoff_``: alpha = 1, the per-epoch ``model_teacher = copy.deepcopy(model_student)``
  of eval_BreastPathQ_SSL_CR.py:515-516 without re-allocating the teacher (parameters and
  floating-point buffers are overwritten bit-exactly; integer buffers are copied with torch).
* ``lookahead_pull_``: the slow-weight update ``p <- a*p + (1-a)*cached; cached <- p`` of
  models/optimiser/RAdam/lookahead.py:96-97 (a = la_alpha).
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Sequence

import torch

from . import _lib
from ._lib import call


def lerp_(dst: Sequence[torch.Tensor], src: Sequence[torch.Tensor], alpha: float,
          write_back: bool = False) -> None:
    dst, src = list(dst), list(src)
    if len(dst) != len(src):
        raise RuntimeError("lerp_: %d destination vs %d source tensors" % (len(dst), len(src)))
    if not dst:
        return
    for d, s in zip(dst, src):
        _lib.require_device(d, "lerp destination")
        if d.dtype != torch.float32 or s.dtype != torch.float32:
            raise RuntimeError("lerp_: only float32 tensors are supported")
        if d.numel() != s.numel() or not d.is_contiguous() or not s.is_contiguous():
            raise RuntimeError("lerp_: tensors must be contiguous and of equal size")
    n = len(dst)
    PtrArr, LLArr = ctypes.c_void_p * n, ctypes.c_longlong * n
    dptr = PtrArr(*[d.data_ptr() for d in dst])
    sptr = PtrArr(*[s.data_ptr() for s in src])
    numel = LLArr(*[d.numel() for d in dst])
    call("b2n_lerp_multi", dptr, sptr, numel, n, float(alpha), 1 if write_back else 0,
         device=dst[0].device)
    _lib.WEIGHT_EPOCH += 1  # parameters changed behind autograd's version counters


def teacher_handoff_(teacher: torch.nn.Module, student: torch.nn.Module) -> None:
    """teacher <- student (alpha = 1), in place, one launch for all float tensors."""
    t_state, s_state = teacher.state_dict(), student.state_dict()
    if list(t_state.keys()) != list(s_state.keys()):
        raise RuntimeError("teacher and student have different state_dict keys")
    fd, fs = [], []
    for k, t in t_state.items():
        s = s_state[k]
        if t.dtype == torch.float32:
            fd.append(t)
            fs.append(s)
        else:
            t.copy_(s)  # num_batches_tracked (int64 scalars)
    lerp_(fd, fs, 1.0)


def lookahead_pull_(params: Iterable[torch.Tensor], cached: Iterable[torch.Tensor],
                    la_alpha: float = 0.5) -> None:
    """p <- la_alpha*p + (1-la_alpha)*cached ; cached <- p."""
    lerp_([p.data for p in params], list(cached), 1.0 - la_alpha, write_back=True)
