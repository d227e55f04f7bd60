"""Masked cross-entropy + argmax + confusion counts as one CUDA pass.

Drop-in for ``nn.CrossEntropyLoss(ignore_index=...)`` as used at
st_water_seg/models/water_seg_model.py:40,103 (mean over non-ignored pixels; an all-ignored
batch gives NaN, which the caller turns into 0 with zero gradients, :104-106) fused with the
``output.argmax(dim=1)`` of :107 and the 3x3 confusion counts the torchmetrics collection of
:46-63 is derived from.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops


class _MaskedCEFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, logits, target, ignore_index, want_pred, confusion):
        if not logits.is_cuda:
            raise RuntimeError("floodplanet_b200 cross-entropy runs on CUDA only (no CPU fallback)")
        if logits.dtype != torch.float32 or target.dtype != torch.int64:
            raise RuntimeError(
                f"cross-entropy expects fp32 logits and int64 targets, got {logits.dtype}/{target.dtype}")
        logits = logits.contiguous()
        target = target.contiguous()
        dev = logits.device
        result = torch.empty(4, dtype=torch.float64, device=dev)
        partials = torch.empty((ops.ce_rows(), 4), dtype=torch.float64, device=dev)
        pred = torch.empty(target.shape, dtype=torch.int64, device=dev) if want_pred else None
        ops.softmax_ce_argmax_fwd(logits, target, ignore_index, result, pred, confusion, partials)
        ctx.save_for_backward(logits, target, result)
        ctx.ignore_index = ignore_index
        loss = result[3].to(torch.float32)
        if pred is None:
            pred = torch.empty(0, dtype=torch.int64, device=dev)
        ctx.mark_non_differentiable(pred, result)
        return loss, pred, result

    @staticmethod
    def backward(ctx, grad_loss, _gp, _gr):
        logits, target, result = ctx.saved_tensors
        dlogits = torch.empty_like(logits)
        go = grad_loss.to(torch.float32).contiguous()
        ops.softmax_ce_bwd(logits, target, ctx.ignore_index, result, go, dlogits)
        return dlogits, None, None, None, None


class MaskedCrossEntropyLoss(nn.Module):
    """``loss = MaskedCrossEntropyLoss(ignore_index)(logits, target)``.

    After each call ``last_pred`` (int64 argmax, first maximum wins), ``last_confusion``
    (int64 [C, C], row = target, col = prediction, ignored pixels excluded) and ``last_result``
    (fp64 [loss_sum, count, n_invalid_targets, mean]) hold the by-products of the same pass.
    """

    def __init__(self, ignore_index: Optional[int] = -100, check_targets: bool = False):
        super().__init__()
        self.ignore_index = -100 if ignore_index is None else int(ignore_index)
        self.check_targets = check_targets
        self.want_pred = True
        self.last_pred: Optional[torch.Tensor] = None
        self.last_confusion: Optional[torch.Tensor] = None
        self.last_result: Optional[torch.Tensor] = None

    def forward(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        ncls = logits.shape[1]
        conf = torch.zeros((ncls, ncls), dtype=torch.int64, device=logits.device)
        loss, pred, result = _MaskedCEFunction.apply(logits, target, self.ignore_index,
                                                     self.want_pred, conf)
        self.last_pred = pred if self.want_pred else None
        self.last_confusion = conf
        self.last_result = result
        if self.check_targets and int(result[2].item()) != 0:
            raise IndexError(f"Target out of bounds for {ncls} classes")  # torch raises likewise
        return loss
