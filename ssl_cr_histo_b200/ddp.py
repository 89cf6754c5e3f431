"""Data-parallel gradient exchange: one flat fp32 arena, one NCCL all-reduce per step.

Replaces ``nn.DataParallel`` (pretrain_BreastPathQ.py:231-233): one process per GPU, weights
resident on every rank (no per-step re-broadcast), each rank runs the step on its batch shard
(BatchNorm statistics stay per-replica exactly as under DataParallel -- the reference has no
SyncBN), then the per-rank mean-loss gradients are averaged with a single all-reduce over
NVLink / NVSwitch.  There is no other collective on the path.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class GradAllReducer:
    """Views every trainable parameter's ``.grad`` into one contiguous buffer so that the
    optimizer keeps working on ``p.grad`` while the exchange is a single collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero_grad(self) -> None:
        """Use instead of optimizer.zero_grad(): keeps .grad aliased into the arena."""
        self.flat.zero_()

    def all_reduce(self, average: bool = True) -> None:
        """Sum gradients over ranks in place with one collective and (``average``) divide by the
        world size; pass ``average=False`` when the optimizer folds 1/world into its step
        (``optim.Adam.grad_scale``)."""
        for p in self.params:  # a backward pass may have re-bound .grad to a fresh tensor
            if p.grad is not None and p.grad.data_ptr() != self._slot(p).data_ptr():
                self._slot(p).copy_(p.grad)
                p.grad = self._slot(p)
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if average:
                self.flat.div_(self.world)

    def _slot(self, p):
        if not hasattr(self, "_slots"):
            self._slots = {}
            off = 0
            for q in self.params:
                self._slots[id(q)] = self.flat[off:off + q.numel()].view_as(q)
                off += q.numel()
        return self._slots[id(p)]


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rows [rank*n/world, (rank+1)*n/world) of dim 0 (the batch axis)."""
    n = t.shape[0]
    if n % world != 0:
        raise RuntimeError("batch of %d does not divide over %d ranks" % (n, world))
    per = n // world
    return t[rank * per:(rank + 1) * per]
