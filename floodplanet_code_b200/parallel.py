"""Data-parallel training of the UNet across the B200s of one box.

The reference is single-GPU (fit.py:86-88 hard-codes devices=1); data parallelism is the new
capability BASELINE.json asks for.  One process per GPU (torchrun), full weight replica,
per-GPU BatchNorm statistics (torch-DDP default semantics, no SyncBN), and ONE exchange step:
the gradient all-reduce.  Backward writes all gradients into a single flat slab in the order it
produces them, so buckets are contiguous slices; each bucket is all-reduced (NCCL, AVG) as
soon as its last layer's wgrad has been enqueued, on a communication stream, while the compute
stream keeps running the remaining backward kernels.

Two transports for the same hook protocol:
  * ``capi`` (default on CUDA): the C ABI's own communicator (``fpb200_nccl_comm_create`` /
    ``fpb200_allreduce_f32`` over an ``ncclComm_t``, include/floodplanet_b200.h) on a stream this
    module owns -- the CTA budget of the collective (``max_ctas``) is then ours to set, which
    matters because it overlaps persistent one-CTA-per-SM convolution kernels;
  * ``torch``: ``torch.distributed.all_reduce`` (NCCL ``AVG``, or gloo ``SUM`` + divide for the
    world_size-2 CPU tests).
torch.distributed is always the control plane (rendezvous, unique-id exchange, barriers).
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


class NcclCommunicator:
    """An ``ncclComm_t`` created through the C ABI (``fpb200_nccl_*``), one per process, on the current
    CUDA device; the unique id travels over torch.distributed (the control plane)."""

    def __init__(self, group=None, max_ctas: int = 0):
        import ctypes as C
        from . import capi
        self._lib = capi.load()
        if self._lib.fpb200_nccl_version() == 0:
            raise RuntimeError("floodplanet_b200: no NCCL runtime found in this process")
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        ident = (C.c_char * 128)()
        if self.rank == 0:
            capi.check(self._lib.fpb200_nccl_unique_id(ident), "nccl_unique_id")
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_char * 128).from_buffer_copy(box[0])
        handle = C.c_void_p()
        capi.check(self._lib.fpb200_nccl_comm_create(C.byref(handle), self.world, self.rank, ident, int(max_ctas)),
                   "nccl_comm_create", world=self.world, rank=self.rank, max_ctas=max_ctas)
        self._comm = handle
        self.max_ctas = int(max_ctas)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.Stream(device=self.device)

    def all_reduce_(self, flat: torch.Tensor, average: bool = True, stream: Optional[torch.cuda.Stream] = None) -> None:
        """In-place fp32 all-reduce of a contiguous 1-D CUDA tensor, enqueued on `stream` (default: the
        communicator's own stream).  The caller orders it against the producer / consumer with events."""
        from . import capi
        if flat.dtype != torch.float32 or not flat.is_cuda or not flat.is_contiguous():
            raise RuntimeError("NcclCommunicator.all_reduce_: contiguous fp32 CUDA tensor expected")
        st = stream if stream is not None else self.stream
        capi.check(self._lib.fpb200_allreduce_f32(self._comm, flat.data_ptr(), flat.numel(), int(average),
                                                  st.cuda_stream), "allreduce_f32", count=flat.numel())

    def destroy(self) -> None:
        if self._comm is not None and self._comm.value:
            torch.cuda.synchronize(self.device)
            self._lib.fpb200_nccl_comm_destroy(self._comm)
        self._comm = None


class BucketedGradAllReduce:
    """Engine hook: all-reduce contiguous slab slices as soon as they are final."""

    def __init__(self, unet: UNet, group=None, bucket_bytes: int = 16 << 20, transport: Optional[str] = None,
                 max_ctas: int = 0):
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
        if transport is None:
            transport = "capi" if self._avg else "torch"
        if transport not in ("capi", "torch"):
            raise ValueError(f"transport {transport!r}: 'capi' or 'torch'")
        self.transport = transport
        self.comm: Optional[NcclCommunicator] = None
        if transport == "capi" and self.world > 1:
            self.comm = NcclCommunicator(group, max_ctas=max_ctas)
        # optional per-bucket timeline: when a list, (start, end, ready_evt, begin_evt, end_evt) per bucket
        self.timeline: Optional[list] = None

    def _launch(self, slab: torch.Tensor, start: int, end: int, side=None) -> None:
        piece = slab[start:end]
        if self.comm is not None:
            # producers (compute stream, and the wgrad side stream when wgrads overlap) -> events ->
            # communication stream -> collective -> event
            compute = torch.cuda.current_stream(slab.device)
            ready = torch.cuda.Event(enable_timing=self.timeline is not None)
            ready.record(compute)
            cs = self.comm.stream
            cs.wait_event(ready)
            if side is not None:
                cs.wait_stream(side)
            t0 = t1 = None
            if self.timeline is not None:
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(cs)
            self.comm.all_reduce_(piece, average=True)
            done = t1 if t1 is not None else torch.cuda.Event()
            done.record(cs)
            if self.timeline is not None:
                self.timeline.append((start, end, ready, t0, t1))
            self._pending.append((done, None))
            return
        if side is not None:
            torch.cuda.current_stream(slab.device).wait_stream(side)   # c10d orders only against the current stream
        if self._avg:
            work = dist.all_reduce(piece, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._pending.append((work, None))
        else:
            work = dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, piece))

    def on_ready(self, slab: torch.Tensor, start: int, end: int, side=None) -> None:
        if self.world == 1:
            return
        if start == 0:
            self._start = 0
            self.buckets_last_step = 0
        if end - self._start >= self.bucket_elems:
            self._launch(slab, self._start, end, side)
            self._start = end
            self.buckets_last_step += 1

    def finish(self, slab: torch.Tensor, total: int) -> None:
        if self.world == 1:
            return
        if total > self._start:
            self._launch(slab, self._start, total)
            self.buckets_last_step += 1
        for work, piece in self._pending:
            if self.comm is not None:
                torch.cuda.current_stream(slab.device).wait_event(work)   # compute waits for the collective
                continue
            work.wait()  # makes the current (compute) stream wait for the collective
            if piece is not None:
                piece.div_(self.world)
        if self.comm is not None:
            slab.record_stream(self.comm.stream)
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
