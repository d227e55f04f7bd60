"""On-device normalise + augment.

The reference normalises and augments every sample on the CPU inside its DataLoader workers
(``BaseDataset.normalize`` st_water_seg/datasets/base_dataset.py:77-113, ``sample_transforms`` :494-530,
``apply_transforms`` :532-555, called from ``__getitem__`` st_water_seg/datasets/floodplanet.py:616-640).
Here the DataLoader only loads and crops; the batch is normalised, flipped and rotated in ONE gather
kernel on the GPU (``csrc/augment.cu``), which can also emit the NHWC bf16 operand of the first
convolution directly.  Sampling keeps the reference's RNG protocol (``np.random``, one coin per active
transform in the order hflip, vflip, rotate; the angle is drawn only when the rotate coin wins), so a
seeded run applies the same transforms the reference would.

There is no CPU path: the ops raise for CPU tensors.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops

FLAG_HFLIP, FLAG_VFLIP, FLAG_ROTATE = 1, 2, 4


def _node(cfg, key):
    """cfg may be an OmegaConf node (attribute access, as in the reference) or a plain dict."""
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


def sample_transforms(transforms) -> List[dict]:
    """`BaseDataset.sample_transforms` (base_dataset.py:494-530): same coins, same order, same dict
    layout ('transform' holds the torchvision function's NAME instead of the function)."""
    active = []
    for name in ("hflip", "vflip"):
        t = _node(transforms, name)
        if _node(t, "active"):
            coin = np.random.rand()
            if coin < _node(t, "likelihood"):
                active.append({"transform": name, "anno": True, "kwargs": {}})
    r = _node(transforms, "rotate")
    if _node(r, "active"):
        coin = np.random.rand()
        if coin < _node(r, "likelihood"):
            rot_angle = np.random.uniform(_node(r, "min_rot_angle"), _node(r, "max_rot_angle"), size=1)[0]
            active.append({"transform": "rotate", "anno": True, "kwargs": {"angle": rot_angle}})
    return active


def _inverse_rotation_matrix(angle: float) -> List[float]:
    """torchvision `_get_inverse_affine_matrix(center=[0, 0], angle=-angle, translate=[0, 0], scale=1,
    shear=[0, 0])` as `transforms.functional.rotate` calls it, in Python doubles."""
    rot = math.radians(-angle)
    a = math.cos(rot)
    b = -math.cos(rot) * math.tan(0.0) - math.sin(rot)
    c = math.sin(rot)
    d = -math.sin(rot) * math.tan(0.0) + math.cos(rot)
    return [d, -b, 0.0, -c, a, 0.0]


def pack_params(active_per_sample: Sequence[Sequence[dict]], h: int, w: int) -> Dict[str, torch.Tensor]:
    """Per-sample transform lists -> the kernel's parameter block (CPU tensors): theta fp32 [N,6] =
    rows of theta^T / [w/2, h/2] formed in fp32 exactly as torchvision's `_gen_affine_grid` does, flags
    int32 [N], and the fp32 linspace base grids."""
    n = len(active_per_sample)
    theta = torch.zeros((n, 6), dtype=torch.float32)
    flags = torch.zeros((n,), dtype=torch.int32)
    scale = torch.tensor([0.5 * w, 0.5 * h], dtype=torch.float32)
    for i, active in enumerate(active_per_sample):
        seen_rotate = False
        for t in active:
            name = t["transform"] if isinstance(t["transform"], str) else t["transform"].__name__
            if seen_rotate:
                raise NotImplementedError("augment: transforms after a rotation are not composed")
            if name == "hflip":
                flags[i] ^= FLAG_HFLIP          # a second flip undoes the first
            elif name == "vflip":
                flags[i] ^= FLAG_VFLIP
            elif name == "rotate":
                m = torch.tensor(_inverse_rotation_matrix(float(t["kwargs"]["angle"])), dtype=torch.float32)
                theta[i] = (m.reshape(1, 2, 3).transpose(1, 2) / scale).reshape(6)
                flags[i] |= FLAG_ROTATE
                seen_rotate = True
            else:
                raise NotImplementedError(f"augment: unknown transform {name!r}")
    return {
        "theta": theta, "flags": flags,
        "xgrid": torch.linspace(-w * 0.5 + 0.5, w * 0.5 + 0.5 - 1, steps=w),
        "ygrid": torch.linspace(-h * 0.5 + 0.5, h * 0.5 + 0.5 - 1, steps=h),
    }


class DeviceAugment:
    """Batch-level drop-in for the per-sample `normalize` + `apply_transforms` of the reference dataset.

    ``transforms``: the `transforms` config node (conf/config.yaml:41-52) or None;
    ``norm_mode``: None | 'local' | 'global' (conf/config.yaml:38); ``global_norm_params``:
    ``{'mean': [C], 'std': [C]}`` for the sensor in use (base_dataset.py:92-94).
    """

    def __init__(self, transforms=None, norm_mode: Optional[str] = None, global_norm_params: Optional[dict] = None):
        if norm_mode not in (None, "local", "global"):
            raise NotImplementedError(f'Normalization mode "{norm_mode}" not implemented.')
        if norm_mode == "global" and global_norm_params is None:
            raise ValueError("norm_mode='global' needs global_norm_params")
        self.transforms = transforms
        self.norm_mode = norm_mode
        self.global_norm_params = global_norm_params
        self._grid_cache: Dict[tuple, tuple] = {}

    def sample(self, n: int) -> List[List[dict]]:
        """One `sample_transforms()` per sample, in batch order."""
        if self.transforms is None:
            return [[] for _ in range(n)]
        return [sample_transforms(self.transforms) for _ in range(n)]

    def normalize_stats(self, image: torch.Tensor):
        n, c = image.shape[:2]
        if self.norm_mode is None:
            return None, None
        if self.norm_mode == "local":
            return ops.plane_mean_std(image)
        mean = torch.as_tensor(np.asarray(self.global_norm_params["mean"], dtype=np.float64), device=image.device)
        std = torch.as_tensor(np.asarray(self.global_norm_params["std"], dtype=np.float64), device=image.device)
        if mean.numel() != c or std.numel() != c:
            raise RuntimeError(f"global norm params have {mean.numel()} channels, image has {c}")
        return mean.expand(n, c).contiguous(), std.expand(n, c).contiguous()

    def __call__(self, batch: Dict[str, torch.Tensor], active: Optional[Sequence[Sequence[dict]]] = None,
                 c_pad: int = 0) -> Dict[str, torch.Tensor]:
        image, target = batch["image"], batch.get("target")
        n, c, h, w = image.shape
        if active is None:
            active = self.sample(n)
        p = pack_params(active, h, w)
        dev = image.device
        key = (h, w, dev)
        if key not in self._grid_cache:
            self._grid_cache[key] = (p["xgrid"].to(dev), p["ygrid"].to(dev))
        xg, yg = self._grid_cache[key]
        mean, std = self.normalize_stats(image)
        out_f32, out_bf16, tgt = ops.augment(image, target, p["theta"].to(dev, non_blocking=True),
                                             p["flags"].to(dev, non_blocking=True), xg, yg, mean, std,
                                             want_f32=True, c_pad=c_pad)
        out = dict(batch)
        out["image"] = out_f32
        if tgt is not None:
            out["target"] = tgt
        if out_bf16 is not None:
            out["image_nhwc_bf16"] = out_bf16
        ones = None
        if mean is None:      # norm_mode None: the reference reports mean 0 / std 1
            mean = torch.zeros((n, c), dtype=torch.float64, device=dev)
            ones = torch.ones((n, c), dtype=torch.float64, device=dev)
        out["mean"] = mean.view(n, c, 1, 1)
        out["std"] = (ones if ones is not None else std).view(n, c, 1, 1)
        out["active_transforms"] = list(active)
        return out
