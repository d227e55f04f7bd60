"""Micro-averaged F1 / Jaccard / Accuracy with ``ignore_index`` from confusion counts.

Stands in for the torchmetrics ``MetricCollection`` of
st_water_seg/models/water_seg_model.py:46-63 (multiclass, ``average='micro'``).  The counts
come for free from the cross-entropy kernel (``MaskedCrossEntropyLoss.last_confusion``), so
the per-step metric update costs no extra pass over the 16.8 M pixels.

With targets equal to ``ignore_index`` removed, micro statistics over C classes are
``tp = trace``, ``fp = fn = total - trace`` and therefore  Accuracy = F1 = tp / total.

Jaccard follows torchmetrics >= 0.11 (the version the reference's ``task="multiclass"`` arguments and
its ``val_MulticlassJaccardIndex`` checkpoint monitor, fit.py:80-85, imply; torchmetrics itself is a
third-party dependency that is not vendored in the reference): per class ``denom_c = colsum_c +
rowsum_c - diag_c``; micro = ``sum(diag) / (sum(denom) - denom[ignore_index])`` when
``0 <= ignore_index < C``.  The ignored class has an empty target row, so ``denom[ignore_index]``
is the number of valid pixels PREDICTED as the ignored class:
  Jaccard = tp / (2*total - tp - colsum[ignore_index])      (= tp / (2*total - tp) without ignore_index).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


class MicroSegmentationMetrics:
    NAMES = ("MulticlassF1Score", "MulticlassJaccardIndex", "MulticlassAccuracy")

    def __init__(self, num_classes: int, ignore_index: Optional[int] = None, prefix: str = ""):
        self.num_classes = num_classes
        self.ignore_index = ignore_index
        self.prefix = prefix
        self._conf: Optional[torch.Tensor] = None

    def clone(self, prefix: str = "") -> "MicroSegmentationMetrics":
        return MicroSegmentationMetrics(self.num_classes, self.ignore_index, prefix)

    # -- accumulation ------------------------------------------------------------------------
    def reset(self) -> None:
        if self._conf is not None:
            self._conf.zero_()

    def update_from_confusion(self, conf: torch.Tensor) -> None:
        # in-place on a persistent device tensor: safe to capture in / replay from a CUDA graph
        if self._conf is None or self._conf.device != conf.device:
            self._conf = torch.zeros_like(conf)
        self._conf.add_(conf)

    def confusion(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """Confusion counts from flat predictions/targets (host-side convenience, not the hot
        path: the training/validation steps use the counts fused into the CE kernel)."""
        c = self.num_classes
        keep = torch.ones_like(target, dtype=torch.bool) if self.ignore_index is None \
            else target != self.ignore_index
        idx = target[keep] * c + pred[keep]
        return torch.bincount(idx, minlength=c * c).view(c, c)

    def update(self, pred: torch.Tensor, target: torch.Tensor) -> None:
        self.update_from_confusion(self.confusion(pred.flatten(), target.flatten()))

    # -- values ------------------------------------------------------------------------------
    def values_from_confusion(self, conf: torch.Tensor) -> Dict[str, torch.Tensor]:
        conf = conf.to(torch.float64)
        tp = conf.diagonal().sum()
        total = conf.sum()
        acc = torch.nan_to_num(tp / total).to(torch.float32)
        denom = 2 * total - tp
        if self.ignore_index is not None and 0 <= self.ignore_index < conf.shape[0]:
            denom = denom - conf[:, self.ignore_index].sum()
        jac = torch.nan_to_num(tp / denom).to(torch.float32)
        p = self.prefix
        return {f"{p}MulticlassF1Score": acc, f"{p}MulticlassJaccardIndex": jac,
                f"{p}MulticlassAccuracy": acc.clone()}

    def compute(self) -> Dict[str, torch.Tensor]:
        if self._conf is None:
            z = torch.zeros(())
            p = self.prefix
            return {f"{p}{n}": z.clone() for n in self.NAMES}
        return self.values_from_confusion(self._conf)

    def forward_from_confusion(self, conf: torch.Tensor) -> Dict[str, torch.Tensor]:
        """torchmetrics ``forward`` semantics: accumulate AND return this batch's values."""
        self.update_from_confusion(conf)
        return self.values_from_confusion(conf)

    def __call__(self, pred: torch.Tensor, target: torch.Tensor) -> Dict[str, torch.Tensor]:
        return self.forward_from_confusion(self.confusion(pred.flatten(), target.flatten()))
