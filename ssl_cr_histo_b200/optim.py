"""Multi-tensor optimizer steps (csrc/optim.cu): drop-in ``torch.optim.Optimizer`` subclasses.

    optimizer = optim.Adam(params, lr=1e-4, weight_decay=1e-4)           # eval_BreastPathQ_SSL_CR.py:481
    optimizer = optim.SGD(params, lr=0.01, momentum=0.9, nesterov=True,   # pretrain_BreastPathQ.py:245
                          weight_decay=1e-4)

Same constructor arguments, ``param_groups`` (so ``MultiStepLR`` and the vendored ``Lookahead``
wrapper keep working), ``zero_grad`` and ``state_dict`` layout as the torch classes they replace,
same arithmetic term by term -- but ``step()`` is one kernel launch per 64 tensors instead of
several per tensor.  ``grad_scale`` (attribute) multiplies every gradient inside the same pass: set
it to 1/world to fold the averaging of a summing all-reduce into the step
(``ddp.GradAllReducer.all_reduce(average=False)``).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import call


def _tables(tensors_by_role):
    n = len(tensors_by_role[0])
    PtrArr = ctypes.c_void_p * n
    out = []
    for role in tensors_by_role:
        out.append(None if role is None else PtrArr(*[t.data_ptr() for t in role]))
    return out, n


def _touch(params) -> None:
    """The kernels write parameters behind autograd's version counters: mark exactly these tensors
    as changed so cached weight packs of *other* modules (a frozen teacher) stay valid."""
    for p in params:
        p._b2n_epoch = getattr(p, "_b2n_epoch", 0) + 1


def _check(p: torch.Tensor, g: torch.Tensor) -> None:
    _lib.require_device(p, "parameter")
    if p.dtype != torch.float32 or g.dtype != torch.float32:
        raise RuntimeError("fused optimizers support float32 parameters and gradients only")
    if g.is_sparse:
        raise RuntimeError("fused optimizers do not support sparse gradients")
    if not p.is_contiguous() or not g.is_contiguous() or g.numel() != p.numel():
        raise RuntimeError("fused optimizers need contiguous parameters and gradients")


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam (L2 weight decay, no amsgrad) in one launch per 64 tensors.

    ``capturable=True`` keeps the step count on the device (one counter per launch group, bumped by
    the launch itself) so that ``step()`` can be captured into a CUDA graph and replayed
    (``graph.GraphedStep``); ``state[p]["step"]`` then counts eager calls only.  Learning rate and
    the other hyper-parameters are baked into a captured launch: re-capture after changing them."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0,
                 capturable=False):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if eps < 0.0:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if weight_decay < 0.0:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = 1.0
        self.capturable = bool(capturable)
        self._step_dev = {}     # capturable: (group index, eager step) -> device counter

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p, p.grad)
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1
                by_step.setdefault(st["step"], []).append(p)
            b1, b2 = group["betas"]
            for step, ps in by_step.items():
                (pt, gt, mt, vt), n = _tables([ps, [p.grad for p in ps],
                                               [self.state[p]["exp_avg"] for p in ps],
                                               [self.state[p]["exp_avg_sq"] for p in ps]])
                numel = (ctypes.c_longlong * n)(*[p.numel() for p in ps])
                counter = None
                if self.capturable:
                    # parameters that have taken the same number of steps share one device counter
                    key = (gi, tuple(id(p) for p in ps))
                    counter = self._step_dev.get(key)
                    if counter is None:
                        counter = self._step_dev[key] = torch.full(
                            (1,), step - 1, device=ps[0].device, dtype=torch.long)
                call("b2n_adam_multi", pt, gt, mt, vt, numel, n, float(group["lr"]), float(b1),
                     float(b2), float(group["eps"]), float(group["weight_decay"]), int(step), counter,
                     float(self.grad_scale), device=ps[0].device)
                _touch(ps)
        return loss


class SGD(torch.optim.Optimizer):
    """torch.optim.SGD (momentum, dampening 0, optional Nesterov, L2 weight decay)."""

    def __init__(self, params, lr=1e-3, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if momentum < 0.0:
            raise ValueError("Invalid momentum value: {}".format(momentum))
        if weight_decay < 0.0:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        if dampening != 0:
            raise NotImplementedError("dampening is not used on the reference's path and not supported")
        if nesterov and momentum <= 0:
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening,
                                      weight_decay=weight_decay, nesterov=nesterov))
        self.grad_scale = 1.0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            fresh, warm = [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p, p.grad)
                st = self.state[p]
                if group["momentum"] != 0 and "momentum_buffer" not in st:
                    st["momentum_buffer"] = torch.empty_like(p, memory_format=torch.preserve_format)
                    fresh.append(p)
                else:
                    warm.append(p)
            for first, ps in ((1, fresh), (0, warm)):
                if not ps:
                    continue
                bufs = [self.state[p]["momentum_buffer"] for p in ps] if group["momentum"] != 0 else None
                (pt, gt, bt), n = _tables([ps, [p.grad for p in ps], bufs])
                numel = (ctypes.c_longlong * n)(*[p.numel() for p in ps])
                call("b2n_sgd_multi", pt, gt, bt, numel, n, float(group["lr"]),
                     float(group["momentum"]), float(group["weight_decay"]),
                     1 if group["nesterov"] else 0, first, float(self.grad_scale),
                     device=ps[0].device)
                _touch(ps)
        return loss
