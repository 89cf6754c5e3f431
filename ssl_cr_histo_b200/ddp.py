"""Data-parallel gradient exchange: one flat fp32 arena, all-reduced over NCCL once per step.

Replaces ``nn.DataParallel`` (pretrain_BreastPathQ.py:231-233): one process per GPU, weights
resident on every rank (no per-step re-broadcast), each rank runs the step on its batch shard
(BatchNorm statistics stay per-replica exactly as under DataParallel -- the reference has no
SyncBN), then the per-rank mean-loss gradients are averaged over NVLink / NVSwitch.  There is no
other collective on the path.

The arena is laid out in *backward order* (heads first, the stem last) and cut into buckets.  The
package's backward kernels write every parameter gradient straight into its arena slot (no
per-parameter ``at::add`` accumulation, no copy), tell the reducer as each one completes, and --
with ``overlap=True`` -- a bucket's all-reduce starts on a side stream as soon as its last gradient
has landed, while the data/weight-gradient kernels of the layers below are still running.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class GradAllReducer:
    """``params``: every trainable parameter of the step, in ``named_parameters()`` order (model
    first, heads after) -- the arena reverses it.  With ``overlap`` the reducer has to know how many
    writers feed each slot per step (three for the trunk under ``TripletNet``'s three passes over
    shared weights, one for its heads): it counts them during the first step (which is reduced
    without overlap) and starts buckets early from the second step on; ``passes`` is only the
    initial guess.  ``sync_initial``: broadcast rank 0's parameters (and,
    via ``sync_buffers``, the floating-point buffers of the given modules) so that replicas that
    were built from different seeds still start identical -- DataParallel's replicate step, done
    once instead of every forward."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, overlap: bool = False,
                 passes: int = 1, bucket_mb: float = 12.0, sync_initial: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = bool(overlap) and self.world > 1
        self.passes = int(passes)
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        # backward order: the last parameter's gradient is complete first
        order = list(reversed(self.params))
        self._slots, self._bucket_of, self.buckets = {}, {}, []
        off, start, members = 0, 0, []
        limit = int(bucket_mb * 2 ** 20 / 4)
        for p in order:
            n = p.numel()
            self._slots[id(p)] = self.flat[off:off + n].view_as(p)
            members.append(p)
            off += n
            if off - start >= limit:
                self._close_bucket(start, off, members)
                start, members = off, []
        if members:
            self._close_bucket(start, off, members)
        for p in self.params:
            slot = self._slots[id(p)]
            p.grad = slot
            # read by the package's backward kernels (trunk.py, heads.py)
            p._b2n_grad_slot = slot
            p._b2n_grad_sink = self if self.overlap else None
        self._pending = [0] * len(self.buckets)
        self._expected = None           # writers per parameter and step, learned in the first step
        self._seen = {}
        self._works = []
        self._stream = torch.cuda.Stream(device=dev) if self.overlap and dev.type == "cuda" else None
        self._arm()
        if sync_initial and self.world > 1:
            self.sync_parameters()

    def _close_bucket(self, start, end, members):
        b = len(self.buckets)
        self.buckets.append((start, end, list(members)))
        for p in members:
            self._bucket_of[id(p)] = b

    def _arm(self):
        for b, (_, _, members) in enumerate(self.buckets):
            if self._expected is None:
                self._pending[b] = 1 << 30          # learning step: launched by all_reduce()
            else:
                counts = [self._expected.get(id(p), 0) for p in members]
                # a member nobody reports (a module outside this package) keeps the bucket manual
                self._pending[b] = sum(counts) if all(counts) else 1 << 30
        self._seen = {}

    # ------------------------------------------------------------------ initial state
    def sync_parameters(self) -> None:
        """Broadcast rank 0's parameter values to every rank (in place)."""
        from . import _lib

        for p in self.params:
            dist.broadcast(p.data, src=0, group=self.group)
        _lib.WEIGHT_EPOCH += 1          # cached weight packs must be rebuilt

    def sync_buffers(self, *modules: torch.nn.Module) -> None:
        """Broadcast rank 0's floating-point buffers (BatchNorm running statistics) once, e.g. after
        loading a checkpoint on rank 0 only.  They are never reduced during training (rank 0's are
        the ones a DataParallel run keeps)."""
        for m in modules:
            for b in m.buffers():
                dist.broadcast(b, src=0, group=self.group)

    # ------------------------------------------------------------------ per step
    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero_grad(self) -> None:
        """Use instead of optimizer.zero_grad(): keeps .grad aliased into the arena."""
        self.flat.zero_()
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self._slots[id(p)].data_ptr():
                p.grad = self._slots[id(p)]
        self._arm()

    def ready(self, p: torch.nn.Parameter) -> None:
        """Called by the backward kernels' host code after a gradient has been enqueued into its
        slot; starts the bucket's all-reduce when its last writer has reported."""
        b = self._bucket_of.get(id(p))
        if b is None or not self.overlap:
            return
        self._seen[id(p)] = self._seen.get(id(p), 0) + 1
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b: int) -> None:
        start, end, _ = self.buckets[b]
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            # a backward pass that runs its weight gradients on a side stream reports gradients from
            # both of its streams: the bucket is complete only when both have caught up
            from . import trunk as _trunk
            for st in _trunk.active_backward_streams():
                self._stream.wait_stream(st)
            with torch.cuda.stream(self._stream):
                work = dist.all_reduce(self.flat[start:end], op=dist.ReduceOp.SUM, group=self.group,
                                       async_op=True)
        else:                           # CPU tensors (gloo): asynchronous on the backend's own thread
            work = dist.all_reduce(self.flat[start:end], op=dist.ReduceOp.SUM, group=self.group,
                                   async_op=True)
        self._works.append(work)
        self._pending[b] = -1           # launched

    def all_reduce(self, average: bool = True) -> None:
        """Finish the exchange: gradients that autograd delivered outside the arena (modules that
        are not this package's) are copied in, every bucket not yet launched is reduced, and the
        launching stream waits for all of it.  ``average=False`` leaves the sum (fold 1/world into
        the optimizer: ``optim.Adam.grad_scale``)."""
        for p in self.params:  # a foreign backward may have re-bound .grad to a fresh tensor
            slot = self._slots[id(p)]
            if p.grad is None:          # optimizer.zero_grad(set_to_none=True) was used
                p.grad = slot
            elif p.grad.data_ptr() != slot.data_ptr():
                if self.overlap and self._pending[self._bucket_of[id(p)]] == -1:
                    raise RuntimeError("a gradient arrived outside the arena after its bucket was "
                                       "reduced; construct GradAllReducer(overlap=False)")
                slot.add_(p.grad)
                p.grad = slot
        if self.overlap and self._expected is None and self._seen:
            self._expected = dict(self._seen)
        if self.world > 1:
            if self.overlap:
                for b in range(len(self.buckets)):
                    if self._pending[b] != -1:
                        self._launch(b)
                for w in self._works:
                    w.wait()            # the current stream waits; the host does not block
                self._works = []
                if self._stream is not None:    # rejoin the side stream (required under graph capture)
                    torch.cuda.current_stream(self.flat.device).wait_stream(self._stream)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if average:
                self.flat.div_(self.world)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rows [rank*n/world, (rank+1)*n/world) of dim 0 (the batch axis)."""
    n = t.shape[0]
    if n % world != 0:
        raise RuntimeError("batch of %d does not divide over %d ranks" % (n, world))
    per = n // world
    return t[rank * per:(rank + 1) * per]
