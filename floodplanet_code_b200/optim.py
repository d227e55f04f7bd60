"""Fused Adam over one flat fp32 parameter slab (reference optimiser:
``optim.Adam(self.parameters(), lr)`` at st_water_seg/models/water_seg_model.py:198-205,
torch defaults betas=(0.9, 0.999), eps=1e-8, no weight decay / amsgrad).

The UNet's parameters are re-homed as views of a single buffer whose layout equals the
engine's gradient slab (reverse-forward order), so one kernel launch updates all 17.27 M
parameters straight from the slab the backward pass (and the data-parallel all-reduce) wrote.
``state_dict`` keys/shapes are unaffected.  Stock ``torch.optim.Adam`` keeps working on the
same module; this class is what the benchmark and the data-parallel trainer use.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .unet import UNet


class FusedAdam:

    def __init__(self, unet: UNet, lr: float, betas=(0.9, 0.999), eps: float = 1e-8):
        self.unet = unet
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.step_count = 0
        self.launches = 0
        engine = unet._engine
        params = dict(unet.named_parameters())
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam needs the model on a CUDA device (no CPU fallback)")
        self.layout, total = engine.grad_layout(params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for name, (off, nel) in self.layout.items():
                p = params[name]
                view = self.flat[off:off + nel].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        # [completed steps, ticket]: the step counter lives on the device so that step() can be
        # captured once and replayed from a CUDA graph
        self.step_state = torch.zeros(2, dtype=torch.int32, device=dev)
        self._first_name = next(iter(self.layout))  # slab offset 0

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.unet.parameters():
            p.grad = None

    def _grad_slab(self) -> Optional[torch.Tensor]:
        """The engine writes every gradient into one slab; recover it from the views."""
        params = dict(self.unet.named_parameters())
        g0 = params[self._first_name].grad
        if g0 is None:
            return None
        base = g0.data_ptr()
        for name, (off, nel) in self.layout.items():
            g = params[name].grad
            if g is None or g.dtype != torch.float32 or g.data_ptr() != base + 4 * off:
                return None
        storage_elems = g0.untyped_storage().nbytes() // 4
        start = g0.storage_offset()
        if start + self.flat.numel() > storage_elems:
            return None
        return torch.as_strided(g0, (self.flat.numel(),), (1,), start)

    def step(self, grad_scale: float = 1.0) -> None:
        self.step_count += 1
        b1, b2 = self.betas
        self.unet._engine.packed.invalidate()  # raw-pointer update: bf16 operand copies are stale
        slab = self._grad_slab()
        if slab is not None:
            ops.adam_step_graphable(self.flat, slab, self.exp_avg, self.exp_avg_sq, self.lr, b1, b2,
                                    self.eps, self.step_state, grad_scale)
            # every master weight just changed: refresh all bf16 operand copies in one launch
            self.launches = 1 + self.unet._engine.packed.refresh_all()
            return
        self.step_count = int(self.step_state[0].item()) + 1
        # gradients came from somewhere else (e.g. accumulated): per-tensor launches
        params = dict(self.unet.named_parameters())
        n = 0
        for name, (off, nel) in self.layout.items():
            g = params[name].grad
            if g is None:
                continue
            ops.adam_step(self.flat[off:off + nel], g.contiguous().view(-1), self.exp_avg[off:off + nel],
                          self.exp_avg_sq[off:off + nel], self.lr, b1, b2, self.eps, self.step_count,
                          grad_scale)
            n += 1
        self.step_state[0] += 1
        self.launches = n
