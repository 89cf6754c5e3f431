"""Fully-connected heads on the libb2n FP32 GEMM kernels (csrc/linear.cu).

``pair_mlp``      : the RSP pair head of models/net.py:56-64 -- cat(E_i, E_j) -> Linear(1024,512) ->
                    ReLU -> Linear(512,256) for the pairs (1,2), (2,3), (1,3), results side by side
                    in one (N,768) tensor.  No concatenation is ever materialised: the first
                    layer's weight is used as its two column halves [Wa | Wb]
                    (cat(A,B) W^T = A Wa^T + B Wb^T, two accumulating GEMMs) and every second-layer
                    GEMM writes straight into its 256-column block of the output.
``pair_mlp_same`` : the same head when E1 = E2 = E3 (models/net.py:88-103, TripletNet_Finetune):
                    the three pair rows are identical, so the hidden layer is computed once.
``mlp2``          : Linear -> ReLU -> Linear  -- ``Classifier.classifier`` (models/net.py:12-15)
``linear``        : single Linear             -- ``FinetuneResNet.classifier`` (models/net.py:110)
All are autograd Functions so the reference's unchanged ``loss.backward()`` reaches the
``nn.Parameter`` leaves; parameter gradients go straight into a flat all-reduce arena when the
parameter has one (``ddp.GradAllReducer``), else into fresh tensors.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call


def _c(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous().float()


class _ParamGrads:
    """Where the parameter gradients of one backward call go.  ``pair(w, b, needed)`` returns
    (dw, db, accumulate) for a Linear's weight and bias (the kernel computes both under one
    accumulate flag): the parameters' arena slots when they have them (accumulate = 1: several
    writers feed a slot) or fresh tensors.  ``result(p)`` is what the autograd Function returns."""

    def __init__(self):
        self.fresh = {}
        self.slots = []

    def pair(self, w: torch.Tensor, b: torch.Tensor, needed: bool):
        if not needed:
            return None, None, 0
        slots = [getattr(p, "_b2n_grad_slot", None) if p.requires_grad else None for p in (w, b)]
        acc = 1 if any(s is not None for s in slots) else 0
        out = []
        for p, s in zip((w, b), slots):
            if s is not None:
                self.slots.append(p)
                out.append(s)
            else:       # (zero-filled when it shares the accumulate flag with a slot)
                t = torch.zeros_like(p) if acc else torch.empty_like(p)
                self.fresh[id(p)] = t
                out.append(t)
        return out[0], out[1], acc

    def result(self, p):
        return self.fresh.get(id(p))

    def done(self):
        for p in self.slots:
            sink = getattr(p, "_b2n_grad_sink", None)
            if sink is not None:
                sink.ready(p)


def _fwd(x, ldx, w, ldw, b, y, ldy, rows, in_f, out_f, relu, acc):
    call("b2n_linear_fwd", x, ldx, w, ldw, b, y, ldy, rows, in_f, out_f, relu, acc)


def _bwd_data(dy, lddy, w, ldw, dx, lddx, mask, rows, in_f, out_f, acc):
    call("b2n_linear_bwd_data", dy, lddy, w, ldw, dx, lddx, mask, rows, in_f, out_f, acc)


def _bwd_weight(dy, lddy, x, ldx, dw, lddw, db, rows, in_f, out_f, acc):
    call("b2n_linear_bwd_weight", dy, lddy, x, ldx, dw, lddw, db, rows, in_f, out_f, acc)


_PAIRS = ((0, 1), (1, 2), (0, 2))          # models/net.py:56-58


class _PairMLPFn(torch.autograd.Function):
    """(E1, E2, E3) -> cat(fc(cat(E1,E2)), fc(cat(E2,E3)), fc(cat(E1,E3))), models/net.py:56-64."""

    @staticmethod
    def forward(ctx, e1, e2, e3, w1p, b1p, w2p, b2p):
        _lib.require_device(e1)
        E = [_c(e1), _c(e2), _c(e3)]
        w1, b1, w2, b2 = _c(w1p), _c(b1p), _c(w2p), _c(b2p)
        n, d = E[0].shape
        hdim, odim = w1.shape[0], w2.shape[0]
        if w1.shape[1] != 2 * d:
            raise RuntimeError("pair MLP expects %d input features, got 2 x %d" % (w1.shape[1], d))
        dev = E[0].device
        h = torch.empty(3, n, hdim, device=dev, dtype=torch.float32)
        y = torch.empty(n, 3 * odim, device=dev, dtype=torch.float32)
        wa, wb = w1, w1[:, d:]                       # column halves of one row-major matrix
        for k, (i, j) in enumerate(_PAIRS):
            _fwd(E[i], d, wa, 2 * d, None, h[k], hdim, n, d, hdim, 0, 0)
            _fwd(E[j], d, wb, 2 * d, b1, h[k], hdim, n, d, hdim, 1, 1)     # += , + bias, ReLU
            _fwd(h[k], hdim, w2, hdim, b2, y[:, k * odim:], 3 * odim, n, hdim, odim, 0, 0)
        ctx.save_for_backward(E[0], E[1], E[2], w1, w2, h)
        ctx.params = (w1p, b1p, w2p, b2p)
        return y

    @staticmethod
    def backward(ctx, dy):
        e1, e2, e3, w1, w2, h = ctx.saved_tensors
        E = [e1, e2, e3]
        w1p, b1p, w2p, b2p = ctx.params
        dy = _c(dy)
        n, d = e1.shape
        hdim, odim = w1.shape[0], w2.shape[0]
        need_e = ctx.needs_input_grad[:3]
        nw1, nb1, nw2, nb2 = ctx.needs_input_grad[3:]
        dev = e1.device
        pg = _ParamGrads()
        dw2, db2, a2 = pg.pair(w2p, b2p, nw2 or nb2)
        dw1, db1, a1 = pg.pair(w1p, b1p, nw1 or nb1)
        dh = torch.empty(3, n, hdim, device=dev, dtype=torch.float32)
        for k in range(3):
            dyk = dy[:, k * odim:]
            if dw2 is not None:
                _bwd_weight(dyk, 3 * odim, h[k], hdim, dw2, hdim, db2, n, hdim, odim, int(a2 or k > 0))
            _bwd_data(dyk, 3 * odim, w2, hdim, dh[k], hdim, h[k], n, hdim, odim, 0)
        if dw1 is not None:
            for k, (i, j) in enumerate(_PAIRS):
                _bwd_weight(dh[k], hdim, E[i], d, dw1, 2 * d, db1, n, d, hdim, int(a1 or k > 0))
                _bwd_weight(dh[k], hdim, E[j], d, dw1[:, d:], 2 * d, None, n, d, hdim, int(a1 or k > 0))
        dE = [None, None, None]
        wa, wb = w1, w1[:, d:]
        for k, (i, j) in enumerate(_PAIRS):
            for idx, wpart in ((i, wa), (j, wb)):
                if not need_e[idx]:
                    continue
                first = dE[idx] is None
                if first:
                    dE[idx] = torch.empty(n, d, device=dev, dtype=torch.float32)
                _bwd_data(dh[k], hdim, wpart, 2 * d, dE[idx], d, None, n, d, hdim, 0 if first else 1)
        pg.done()
        return (dE[0], dE[1], dE[2], pg.result(w1p) if nw1 else None, pg.result(b1p) if nb1 else None,
                pg.result(w2p) if nw2 else None, pg.result(b2p) if nb2 else None)


class _PairMLPSameFn(torch.autograd.Function):
    """E -> cat(f, f, f) with f = fc(cat(E, E)): models/net.py:92-103 when the three trunk passes
    saw the same input.  One hidden-layer evaluation; the gradients are those of the three-pair
    graph (the three upstream blocks are summed)."""

    @staticmethod
    def forward(ctx, e, w1p, b1p, w2p, b2p):
        _lib.require_device(e)
        e, w1, b1, w2, b2 = _c(e), _c(w1p), _c(b1p), _c(w2p), _c(b2p)
        n, d = e.shape
        hdim, odim = w1.shape[0], w2.shape[0]
        if w1.shape[1] != 2 * d:
            raise RuntimeError("pair MLP expects %d input features, got 2 x %d" % (w1.shape[1], d))
        dev = e.device
        h = torch.empty(n, hdim, device=dev, dtype=torch.float32)
        y = torch.empty(n, 3 * odim, device=dev, dtype=torch.float32)
        _fwd(e, d, w1, 2 * d, None, h, hdim, n, d, hdim, 0, 0)
        _fwd(e, d, w1[:, d:], 2 * d, b1, h, hdim, n, d, hdim, 1, 1)
        _fwd(h, hdim, w2, hdim, b2, y, 3 * odim, n, hdim, odim, 0, 0)          # block 0 ...
        call("b2n_cols_replicate", y, 3 * odim, n, odim, 3)                     # ... = blocks 1, 2
        ctx.save_for_backward(e, w1, w2, h)
        ctx.params = (w1p, b1p, w2p, b2p)
        return y

    @staticmethod
    def backward(ctx, dy):
        e, w1, w2, h = ctx.saved_tensors
        w1p, b1p, w2p, b2p = ctx.params
        dy = _c(dy)
        n, d = e.shape
        hdim, odim = w1.shape[0], w2.shape[0]
        ne, nw1, nb1, nw2, nb2 = ctx.needs_input_grad
        dev = e.device
        pg = _ParamGrads()
        dw2, db2, a2 = pg.pair(w2p, b2p, nw2 or nb2)
        dw1, db1, a1 = pg.pair(w1p, b1p, nw1 or nb1)
        dh = torch.empty(n, hdim, device=dev, dtype=torch.float32)
        df = torch.empty(n, odim, device=dev, dtype=torch.float32)
        call("b2n_cols_sum", dy, 3 * odim, df, n, odim, 3)     # the three pair rows share f
        if dw2 is not None:
            _bwd_weight(df, odim, h, hdim, dw2, hdim, db2, n, hdim, odim, a2)
        _bwd_data(df, odim, w2, hdim, dh, hdim, h, n, hdim, odim, 0)
        if dw1 is not None:
            # each of the three pair rows contributes dh_k^T E to both column halves and
            # colsum(dh_k) to the bias; the rows are identical, so dh already holds the sum
            _bwd_weight(dh, hdim, e, d, dw1, 2 * d, db1, n, d, hdim, a1)
            _bwd_weight(dh, hdim, e, d, dw1[:, d:], 2 * d, None, n, d, hdim, a1)
        de = None
        if ne:
            de = torch.empty(n, d, device=dev, dtype=torch.float32)
            _bwd_data(dh, hdim, w1, 2 * d, de, d, None, n, d, hdim, 0)
            _bwd_data(dh, hdim, w1[:, d:], 2 * d, de, d, None, n, d, hdim, 1)
        pg.done()
        return (de, pg.result(w1p) if nw1 else None, pg.result(b1p) if nb1 else None,
                pg.result(w2p) if nw2 else None, pg.result(b2p) if nb2 else None)


class _MLP2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1p, b1p, w2p, b2p):
        _lib.require_device(x)
        x, w1, b1, w2, b2 = _c(x), _c(w1p), _c(b1p), _c(w2p), _c(b2p)
        n, k = x.shape
        hdim, odim = w1.shape[0], w2.shape[0]
        h = torch.empty(n, hdim, device=x.device, dtype=torch.float32)
        y = torch.empty(n, odim, device=x.device, dtype=torch.float32)
        _fwd(x, k, w1, k, b1, h, hdim, n, k, hdim, 1, 0)
        _fwd(h, hdim, w2, hdim, b2, y, odim, n, hdim, odim, 0, 0)
        ctx.save_for_backward(x, w1, w2, h)
        ctx.params = (w1p, b1p, w2p, b2p)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2, h = ctx.saved_tensors
        w1p, b1p, w2p, b2p = ctx.params
        dy = _c(dy)
        n, k = x.shape
        hdim, odim = w1.shape[0], w2.shape[0]
        nx, nw1, nb1, nw2, nb2 = ctx.needs_input_grad
        dev = x.device
        pg = _ParamGrads()
        dx = None
        dw2, db2, a2 = pg.pair(w2p, b2p, nw2 or nb2)
        if dw2 is not None:
            _bwd_weight(dy, odim, h, hdim, dw2, hdim, db2, n, hdim, odim, a2)
        if nx or nw1 or nb1:
            dh = torch.empty(n, hdim, device=dev)
            _bwd_data(dy, odim, w2, hdim, dh, hdim, h, n, hdim, odim, 0)
            dw1, db1, a1 = pg.pair(w1p, b1p, nw1 or nb1)
            if dw1 is not None:
                _bwd_weight(dh, hdim, x, k, dw1, k, db1, n, k, hdim, a1)
            if nx:
                dx = torch.empty(n, k, device=dev)
                _bwd_data(dh, hdim, w1, k, dx, k, None, n, k, hdim, 0)
        pg.done()
        return (dx, pg.result(w1p) if nw1 else None, pg.result(b1p) if nb1 else None,
                pg.result(w2p) if nw2 else None, pg.result(b2p) if nb2 else None)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wp, bp):
        _lib.require_device(x)
        x, w, b = _c(x), _c(wp), _c(bp)
        n, k = x.shape
        o = w.shape[0]
        y = torch.empty(n, o, device=x.device, dtype=torch.float32)
        _fwd(x, k, w, k, b, y, o, n, k, o, 0, 0)
        ctx.save_for_backward(x, w)
        ctx.params = (wp, bp)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        wp, bp = ctx.params
        dy = _c(dy)
        n, k = x.shape
        o = w.shape[0]
        nx, nw, nb = ctx.needs_input_grad
        pg = _ParamGrads()
        dx = None
        dw, db, a = pg.pair(wp, bp, nw or nb)
        if dw is not None:
            _bwd_weight(dy, o, x, k, dw, k, db, n, k, o, a)
        if nx:
            dx = torch.empty(n, k, device=x.device)
            _bwd_data(dy, o, w, k, dx, k, None, n, k, o, 0)
        pg.done()
        return dx, pg.result(wp) if nw else None, pg.result(bp) if nb else None


def pair_mlp(e1, e2, e3, lin1, lin2):
    """(N,768) pair features of models/net.py:56-64 from the three (N,512) trunk outputs."""
    return _PairMLPFn.apply(e1, e2, e3, lin1.weight, lin1.bias, lin2.weight, lin2.bias)


def pair_mlp_same(e, lin1, lin2):
    """``pair_mlp(e, e, e, ...)`` computed once (TripletNet_Finetune)."""
    return _PairMLPSameFn.apply(e, lin1.weight, lin1.bias, lin2.weight, lin2.bias)


def mlp2(x, lin1, lin2):
    """x -> lin2(relu(lin1(x))) with ``nn.Linear`` parameter containers lin1 / lin2."""
    return _MLP2Fn.apply(x, lin1.weight, lin1.bias, lin2.weight, lin2.bias)


def linear(x, lin):
    return _LinearFn.apply(x, lin.weight, lin.bias)
