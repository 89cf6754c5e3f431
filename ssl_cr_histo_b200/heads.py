"""Fully-connected heads on the libb2n FP32 GEMM kernels (csrc/linear.cu).

``mlp2``  : Linear -> ReLU -> Linear  -- the pair-MLP ``fc`` (models/net.py:36-37) and
            ``Classifier.classifier`` (models/net.py:12-15)
``linear``: single Linear             -- ``FinetuneResNet.classifier`` (models/net.py:110)
Both are autograd Functions so the reference's unchanged ``loss.backward()`` reaches the
``nn.Parameter`` leaves.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call


def _c(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous().float()


class _MLP2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        _lib.require_device(x)
        x, w1, b1, w2, b2 = _c(x), _c(w1), _c(b1), _c(w2), _c(b2)
        n, k = x.shape
        hdim, odim = w1.shape[0], w2.shape[0]
        h = torch.empty(n, hdim, device=x.device, dtype=torch.float32)
        y = torch.empty(n, odim, device=x.device, dtype=torch.float32)
        call("b2n_linear_fwd", x, k, w1, k, b1, h, hdim, n, k, hdim, 1, 0)
        call("b2n_linear_fwd", h, hdim, w2, hdim, b2, y, odim, n, hdim, odim, 0, 0)
        ctx.save_for_backward(x, w1, w2, h)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2, h = ctx.saved_tensors
        dy = _c(dy)
        n, k = x.shape
        hdim, odim = w1.shape[0], w2.shape[0]
        nx, nw1, nb1, nw2, nb2 = ctx.needs_input_grad
        dev = x.device
        dw2 = db2 = dw1 = db1 = dx = None
        if nw2 or nb2:
            dw2 = torch.empty(odim, hdim, device=dev)
            db2 = torch.empty(odim, device=dev)
            call("b2n_linear_bwd_weight", dy, odim, h, hdim, dw2, hdim, db2, n, hdim, odim, 0)
        if nx or nw1 or nb1:
            dh = torch.empty(n, hdim, device=dev)
            call("b2n_linear_bwd_data", dy, odim, w2, hdim, dh, hdim, h, n, hdim, odim, 0)
            if nw1 or nb1:
                dw1 = torch.empty(hdim, k, device=dev)
                db1 = torch.empty(hdim, device=dev)
                call("b2n_linear_bwd_weight", dh, hdim, x, k, dw1, k, db1, n, k, hdim, 0)
            if nx:
                dx = torch.empty(n, k, device=dev)
                call("b2n_linear_bwd_data", dh, hdim, w1, k, dx, k, None, n, k, hdim, 0)
        return dx, dw1 if nw1 else None, db1 if nb1 else None, dw2 if nw2 else None, \
            db2 if nb2 else None


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        _lib.require_device(x)
        x, w, b = _c(x), _c(w), _c(b)
        n, k = x.shape
        o = w.shape[0]
        y = torch.empty(n, o, device=x.device, dtype=torch.float32)
        call("b2n_linear_fwd", x, k, w, k, b, y, o, n, k, o, 0, 0)
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _c(dy)
        n, k = x.shape
        o = w.shape[0]
        nx, nw, nb = ctx.needs_input_grad
        dx = dw = db = None
        if nw or nb:
            dw = torch.empty(o, k, device=x.device)
            db = torch.empty(o, device=x.device)
            call("b2n_linear_bwd_weight", dy, o, x, k, dw, k, db, n, k, o, 0)
        if nx:
            dx = torch.empty(n, k, device=x.device)
            call("b2n_linear_bwd_data", dy, o, w, k, dx, k, None, n, k, o, 0)
        return dx, dw if nw else None, db if nb else None


def mlp2(x, lin1, lin2):
    """x -> lin2(relu(lin1(x))) with ``nn.Linear`` parameter containers lin1 / lin2."""
    return _MLP2Fn.apply(x, lin1.weight, lin1.bias, lin2.weight, lin2.bias)


def linear(x, lin):
    return _LinearFn.apply(x, lin.weight, lin.bias)
