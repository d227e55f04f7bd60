"""Whole-step CUDA graph: forward + masked CE + backward + fused Adam captured once and replayed.

One training step is ~200 kernel launches, ~80 of them a few microseconds long (BatchNorm
finalises, weight re-packs): issued one by one from Python the GPU outruns the host.  The step
has static shapes and no host synchronisation, so it is captured into a CUDA graph and replayed
with one launch.  Inputs live in static device buffers (`image`, `target`) that the caller
fills -- e.g. directly with the host->device copy of the next batch.
"""
from __future__ import annotations

from typing import Dict

import torch

from .engine import note_raw_parameter_write
from .optim import FusedAdam


class GraphedTrainStep:
    """``step = GraphedTrainStep(lightning_module, fused_adam, example_batch)``;
    fill ``step.batch[...]`` and call ``step.replay()`` -> static loss tensor."""

    def __init__(self, module, optimizer: FusedAdam, example_batch: Dict[str, torch.Tensor],
                 warmup_steps: int = 3):
        self.module = module
        self.optimizer = optimizer
        self.batch = {k: v.clone() for k, v in example_batch.items() if isinstance(v, torch.Tensor)}
        engine = module.model._engine
        if engine.grad_ready_hook is not None:
            raise RuntimeError("GraphedTrainStep: capture with a gradient all-reduce hook is not supported")
        # one captured stream: the wgrad side stream's allocator bookkeeping (record_stream) does not belong in a
        # capture, and inside a graph the kernels are already free of launch gaps
        engine.overlap_wgrad = False
        self.fwd_launches = 0
        self.bwd_launches = 0
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up on a side stream (allocator, lazy init)
            for i in range(warmup_steps):
                self._eager_step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        engine.packed.always_repack = True
        self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph):
                self.loss = self._eager_step(0)
        finally:
            engine.packed.always_repack = False
        # gradients stay views of the captured slab; parameters were updated in place

    def _eager_step(self, i: int) -> torch.Tensor:
        self.optimizer.zero_grad()
        loss = self.module.training_step(self.batch, i)
        self.fwd_launches = self.module.model._engine.launches
        loss.backward()
        self.bwd_launches = self.module.model._engine.launches
        self.optimizer.step()
        return loss.detach()

    @property
    def kernel_launches(self) -> int:
        """kernels of this repo inside one replay (forward + CE fwd/bwd + backward + Adam)."""
        return self.fwd_launches + self.bwd_launches + 3 + 1

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        # the replay updated the master weights, BatchNorm affine parameters and running statistics
        # through raw pointers: no tensor version moved, so announce it (host-only, free) -- otherwise
        # an eval / validation forward after replays would reuse stale packed weights and folded
        # BatchNorm coefficients
        note_raw_parameter_write()
        return self.loss
