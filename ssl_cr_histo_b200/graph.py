"""Whole-step CUDA-graph capture: forward, loss, backward and optimizer step replayed as one graph.

A consistency step is ~250 kernel launches issued through ctypes and autograd; at BASELINE batch
sizes the GPU hides that, at small per-rank batches (or on a busy host) it does not.  Every kernel
of this package is launched on the caller's stream with caller-owned memory and no host
synchronisation, so a step can be captured once and replayed:

    step = GraphedStep(train_step, (inputs_x, targets_x, inputs_u_w, inputs_u_s))
    for batch in loader:
        loss = step(*batch)              # copies the batch into the static inputs, replays

Requirements on ``fn``: static shapes; no host read-back (``.item()``) inside; the optimizer must be
capturable (``optim.Adam(capturable=True)``, ``optim.SGD`` after its first eager step -- or
torch.optim's own ``capturable=True``); gradients are zeroed with ``set_to_none=True`` or through
``ddp.GradAllReducer.zero_grad()`` inside ``fn``.  The warm-up calls are real training steps.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        self.static_in = [torch.empty_like(t) for t in example_inputs]
        for d, s in zip(self.static_in, example_inputs):
            d.copy_(s)
        dev = self.static_in[0].device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):         # warm up off the default stream (allocator, packs,
            for _ in range(max(warmup, 1)):   # lazily built optimizer state, function attributes)
                fn(*self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def load(self, *inputs: torch.Tensor) -> None:
        """Copy a batch (device tensors, or pinned host tensors: asynchronous H2D) into the graph's
        static inputs on the current stream."""
        for d, s in zip(self.static_in, inputs):
            d.copy_(s, non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.static_out

    def __call__(self, *inputs: torch.Tensor):
        self.load(*inputs)
        return self.replay()
