"""Data-parallel training of the UNet across the B200s of one box.

The reference is single-GPU (fit.py:86-88 hard-codes devices=1); data parallelism is the new
capability BASELINE.json asks for.  One process per GPU (torchrun), full weight replica,
per-GPU BatchNorm statistics (torch-DDP default semantics, no SyncBN), and ONE exchange step:
the gradient all-reduce.  Backward writes all gradients into a single flat slab in the order it
produces them, so buckets are contiguous slices; each bucket is all-reduced (NCCL, AVG) as
soon as its last layer's wgrad has been enqueued, on NCCL's own stream, while the compute
stream keeps running the remaining backward kernels.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from .engine import note_raw_parameter_write
from .unet import UNet


def init_distributed(backend: Optional[str] = None) -> tuple:
    """(rank, world_size, local_rank) from the torchrun environment; no-op for one process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced partition of n_items work units (scene tiles, samples) over ranks."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class BucketedGradAllReduce:
    """Engine hook: all-reduce contiguous slab slices as soon as they are final."""

    def __init__(self, unet: UNet, group=None, bucket_bytes: int = 16 << 20):
        self.group = group
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._pending: List = []
        self._start = 0
        self.buckets_last_step = 0
        unet._engine.grad_ready_hook = self.on_ready
        unet._engine.grad_done_hook = self.finish
        # NCCL has ReduceOp.AVG; gloo (CPU tests) does not
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"

    def _launch(self, slab: torch.Tensor, start: int, end: int) -> None:
        piece = slab[start:end]
        if self._avg:
            work = dist.all_reduce(piece, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._pending.append((work, None))
        else:
            work = dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, piece))

    def on_ready(self, slab: torch.Tensor, start: int, end: int) -> None:
        if self.world == 1:
            return
        if start == 0:
            self._start = 0
            self.buckets_last_step = 0
        if end - self._start >= self.bucket_elems:
            self._launch(slab, self._start, end)
            self._start = end
            self.buckets_last_step += 1

    def finish(self, slab: torch.Tensor, total: int) -> None:
        if self.world == 1:
            return
        if total > self._start:
            self._launch(slab, self._start, total)
            self.buckets_last_step += 1
        for work, piece in self._pending:
            work.wait()  # makes the current (compute) stream wait for the collective
            if piece is not None:
                piece.div_(self.world)
        self._pending.clear()
        self._start = 0


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every replica start from rank `src`'s weights and BatchNorm buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
    # writes through .data do not move Parameter._version: announce them so that packed bf16 weight
    # copies and folded eval-mode BatchNorm coefficients made by an earlier forward are rebuilt
    note_raw_parameter_write()
