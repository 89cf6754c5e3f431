"""Fused loss kernels (csrc/loss_lerp.cu) as autograd-aware functions.

Class targets must lie in [0, C): an out-of-range label (``ignore_index`` is not supported -- the
reference never uses it) is not dereferenced and turns the loss and its gradient into NaN.

The reference's loops call torch losses themselves (pretrain_BreastPathQ.py:56,
eval_BreastPathQ_SSL_CR.py:92-95, eval_Kather_SSL_CR.py:87-93); those keep working with the
drop-in modules.  These functions are the opt-in single-launch versions: one pass over the
logits produces the loss values, d(total)/d(logits) and the decisions (argmax / pseudo labels).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mode, lambda_u, logits_x, targets, logits_u_w, logits_u_s):
        _lib.require_device(logits_x, "logits")
        dev = logits_x.device
        if logits_x.dim() != 2:
            raise RuntimeError("logits_x must be (rows, classes), got %s" % (tuple(logits_x.shape),))
        lx = logits_x.contiguous().float()
        rows_x, C = lx.shape
        losses = torch.empty(3, device=dev)
        dlx = torch.empty_like(lx)
        ti = tf = lw = ls = dlu = amax = pseudo = None
        rows_u = 0
        if mode == 1:
            tf = targets.contiguous().float().view(-1)
            if tf.numel() != lx.numel():
                raise RuntimeError("MSE targets must have as many elements as logits_x")
        else:
            ti = targets.contiguous().long().view(-1)
            if ti.numel() != rows_x:
                raise RuntimeError("expected %d class targets, got %d" % (rows_x, ti.numel()))
            amax = torch.empty(rows_x, device=dev, dtype=torch.long)
        if mode != 0:
            if logits_u_w.shape != logits_u_s.shape or logits_u_s.dim() != 2 or \
                    logits_u_s.shape[1] != C:
                raise RuntimeError("logits_u_w %s and logits_u_s %s must both be (rows_u, %d)"
                                   % (tuple(logits_u_w.shape), tuple(logits_u_s.shape), C))
            lw = logits_u_w.detach().contiguous().float()
            ls = logits_u_s.contiguous().float()
            rows_u = ls.shape[0]
            dlu = torch.empty_like(ls)
            if mode == 2:
                pseudo = torch.empty(rows_u, device=dev, dtype=torch.long)
        call("b2n_fused_loss", mode, lx, ti, tf, lw, ls, rows_x, rows_u, C, float(lambda_u), losses,
             dlx, dlu, amax, pseudo)
        ctx.save_for_backward(dlx, dlu if dlu is not None else dlx)
        ctx.has_u = dlu is not None
        ctx.mark_non_differentiable(*(t for t in (amax, pseudo) if t is not None))
        total = losses[2] if mode != 0 else losses[0]
        empty = torch.empty(0, device=dev, dtype=torch.long)
        return (total, losses.detach().clone(), amax if amax is not None else empty,
                pseudo if pseudo is not None else empty)

    @staticmethod
    def backward(ctx, g_total, _g_losses, _g_amax, _g_pseudo):
        dlx, dlu = ctx.saved_tensors
        gx = dlx * g_total
        gu = dlu * g_total if ctx.has_u else None
        return None, None, gx, None, None, gu


def cross_entropy(logits, target):
    """mean softmax-CE + argmax in one launch -> (loss, pred).  nn.CrossEntropyLoss()(output,
    target) and torch.argmax(output, 1) of pretrain_BreastPathQ.py:56,66."""
    total, _, amax, _ = _FusedLoss.apply(0, 0.0, logits, target, None, None)
    return total, amax


def consistency_mse(logits_x, targets_x, logits_u_w, logits_u_s, lambda_u=1.0):
    """eval_BreastPathQ_SSL_CR.py:92-95 -> (final_loss, [sup, cons, final])."""
    total, parts, _, _ = _FusedLoss.apply(1, lambda_u, logits_x, targets_x, logits_u_w, logits_u_s)
    return total, parts


def consistency_ce(logits_x, targets_x, logits_u_w, logits_u_s, lambda_u=1.0):
    """eval_Kather_SSL_CR.py:87-93 -> (final_loss, [sup, cons, final], pred_x, pseudo_labels)."""
    total, parts, amax, pseudo = _FusedLoss.apply(2, lambda_u, logits_x, targets_x, logits_u_w,
                                                  logits_u_s)
    return total, parts, amax, pseudo
