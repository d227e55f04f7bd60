"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-sample normalise + augment
step, the checker for ``floodplanet_code_b200.augment``.

Only ``tests/`` may import this module; the product path never does.

Follows, line by line:

* ``BaseDataset.normalize``          st_water_seg/datasets/base_dataset.py:77-113
* ``BaseDataset.sample_transforms``  st_water_seg/datasets/base_dataset.py:494-530
* ``BaseDataset.apply_transforms``   st_water_seg/datasets/base_dataset.py:532-555
* call order in ``__getitem__``      st_water_seg/datasets/floodplanet.py:616-640
  (normalize -> buffer -> hflip -> vflip -> rotate; image ``.float()``, target ``.long()``)

The arithmetic of the three transforms lives in a third-party dependency, torchvision
(``torchvision.transforms.functional.{hflip,vflip,rotate}``; the reference pins torchvision 0.11.3,
``environment.yml``; this container has 0.26.0).  ``rotate`` is called with its defaults
(``base_dataset.py:519-526``): nearest interpolation, ``expand=False``, centre = image centre,
fill 0.  Its published algorithm is restated here in numpy:

1. inverse affine matrix of a pure rotation about the centre, computed in Python doubles
   (``_get_inverse_affine_matrix(center=[0,0], angle=-angle, translate=[0,0], scale=1, shear=[0,0])``);
2. ``_gen_affine_grid``: fp32 ``theta^T / [w/2, h/2]`` applied to the base grid
   ``linspace(-w/2 + .5, w/2 - .5, w)`` x ``linspace(-h/2 + .5, h/2 - .5, h)`` by ``bmm``;
   the CPU ``bmm`` evaluates ``fma(y, t_y, x * t_x) + t_0`` in fp32 (pinned empirically: every one of
   262 144 grid values bit-equal for all tested angles, other association orders differ);
3. ``grid_sample(mode='nearest', padding_mode='zeros', align_corners=False)`` on CPU:
   ``ix = (gx + 1) * (W / 2) - 0.5`` in fp32, ``nearbyint`` (ties to even), zero outside the image.

**Pinned**: ``tests/test_augment_cpu.py`` checks the restated index map against torchvision itself
(when importable) for many angles / sizes, and ``tests/golden/augment.pt`` holds outputs of the
reference's OWN ``BaseDataset.normalize / sample_transforms / apply_transforms`` (imported by file path
with stand-ins for tifffile / pytorch_lightning, see ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------
# normalisation (base_dataset.py:77-113)
# ---------------------------------------------------------------------------------------------
def normalize(image: np.ndarray, norm_mode: Optional[str], global_params: Optional[dict] = None
              ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """image [C,H,W] float -> (normalised image, mean [C,1,1], std [C,1,1]); `image -= mean; image /= std`."""
    image = np.array(image, copy=True)
    if norm_mode == "global":
        mean = np.asarray(global_params["mean"])[:, None, None]
        std = np.asarray(global_params["std"])[:, None, None]
    elif norm_mode == "local":
        flat = image.reshape(image.shape[0], image.shape[1] * image.shape[2])
        mean = flat.mean(axis=1)[:, None, None]
        std = flat.std(axis=1)[:, None, None]          # population std (ddof = 0)
    elif norm_mode is None:
        mean = np.zeros([image.shape[0], 1, 1], dtype=image.dtype)
        std = np.ones([image.shape[0], 1, 1], dtype=image.dtype)
    else:
        raise NotImplementedError(f'Normalization mode "{norm_mode}" not implemented.')
    # numpy in-place ops: evaluated in the wider of (image, statistics) dtypes, rounded to the image's
    image -= mean
    image /= std
    return image, mean, std


# ---------------------------------------------------------------------------------------------
# sampling (base_dataset.py:494-530): one np.random.rand() per ACTIVE transform, in the order
# hflip, vflip, rotate; rotate draws its angle with np.random.uniform only when its coin wins
# ---------------------------------------------------------------------------------------------
def sample_transforms(cfg: dict, rng=np.random) -> List[dict]:
    active = []
    if cfg["hflip"]["active"]:
        if rng.rand() < cfg["hflip"]["likelihood"]:
            active.append({"transform": "hflip", "anno": True, "kwargs": {}})
    if cfg["vflip"]["active"]:
        if rng.rand() < cfg["vflip"]["likelihood"]:
            active.append({"transform": "vflip", "anno": True, "kwargs": {}})
    if cfg["rotate"]["active"]:
        if rng.rand() < cfg["rotate"]["likelihood"]:
            angle = rng.uniform(cfg["rotate"]["min_rot_angle"], cfg["rotate"]["max_rot_angle"], size=1)[0]
            active.append({"transform": "rotate", "anno": True, "kwargs": {"angle": angle}})
    return active


# ---------------------------------------------------------------------------------------------
# torchvision.transforms.functional.rotate, nearest / expand=False / centre / fill 0
# ---------------------------------------------------------------------------------------------
def inverse_rotation_matrix(angle: float) -> List[float]:
    """`_get_inverse_affine_matrix([0, 0], -angle, [0, 0], 1.0, [0, 0])` in Python doubles."""
    rot = math.radians(-angle)
    sx = sy = math.radians(0.0)
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m = [x / 1.0 for x in m]
    m[2] += m[0] * (-0.0 - 0.0) + m[1] * (-0.0 - 0.0)
    m[5] += m[3] * (-0.0 - 0.0) + m[4] * (-0.0 - 0.0)
    m[2] += 0.0
    m[5] += 0.0
    return m


def rescaled_theta(angle: float, h: int, w: int) -> np.ndarray:
    """fp32 [3, 2]: theta^T / [0.5 w, 0.5 h] exactly as `_gen_affine_grid` forms it."""
    theta = torch.tensor(inverse_rotation_matrix(angle), dtype=torch.float32).reshape(1, 2, 3)
    rt = theta.transpose(1, 2) / torch.tensor([0.5 * w, 0.5 * h], dtype=torch.float32)
    return rt[0].numpy().copy()


def base_grids(h: int, w: int) -> Tuple[np.ndarray, np.ndarray]:
    xg = torch.linspace(-w * 0.5 + 0.5, w * 0.5 + 0.5 - 1, steps=w).numpy().copy()
    yg = torch.linspace(-h * 0.5 + 0.5, h * 0.5 + 0.5 - 1, steps=h).numpy().copy()
    return xg, yg


def _fma32(a, b, c):
    # exact: the product of two fp32 values and the sum fit a double with room to spare for
    # one correctly rounded fp32 result in every case that matters here (|values| < 2)
    return (np.float64(a) * np.float64(b) + np.float64(c)).astype(np.float32)


def rotate_nearest_index_map(angle: float, h: int, w: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Source (row, col) read by every output pixel and whether it lies inside the image."""
    rt = rescaled_theta(angle, h, w)
    xg, yg = base_grids(h, w)
    X = np.broadcast_to(xg[None, :], (h, w)).astype(np.float32)
    Y = np.broadcast_to(yg[:, None], (h, w)).astype(np.float32)
    g = []
    for c in range(2):
        p = (X * rt[0, c]).astype(np.float32)
        g.append((_fma32(Y, rt[1, c], p) + rt[2, c]).astype(np.float32))
    ix = ((g[0] + np.float32(1)) * np.float32(w / 2) - np.float32(0.5)).astype(np.float32)
    iy = ((g[1] + np.float32(1)) * np.float32(h / 2) - np.float32(0.5)).astype(np.float32)
    xn, yn = np.rint(ix), np.rint(iy)
    ok = (xn >= 0) & (xn <= w - 1) & (yn >= 0) & (yn <= h - 1)
    return (np.where(ok, yn, 0).astype(np.int64), np.where(ok, xn, 0).astype(np.int64), ok)


def apply_transforms(image, active: Sequence[dict], is_anno: bool) -> torch.Tensor:
    """numpy/torch [..., H, W] -> torch tensor; every sampled transform has anno=True, so image and
    annotation go through the same geometric chain (base_dataset.py:532-555)."""
    t = torch.as_tensor(np.asarray(image)) if not isinstance(image, torch.Tensor) else image
    for tr in active:
        if is_anno and not tr["anno"]:
            continue
        if tr["transform"] == "hflip":
            t = t.flip(-1)
        elif tr["transform"] == "vflip":
            t = t.flip(-2)
        elif tr["transform"] == "rotate":
            h, w = t.shape[-2], t.shape[-1]
            sy, sx, ok = rotate_nearest_index_map(float(tr["kwargs"]["angle"]), h, w)
            src = t[..., torch.from_numpy(sy), torch.from_numpy(sx)]
            t = torch.where(torch.from_numpy(ok), src, torch.zeros((), dtype=t.dtype))
        else:
            raise NotImplementedError(tr["transform"])
    return t


def augment_sample(image: np.ndarray, target: np.ndarray, active: Sequence[dict], norm_mode: Optional[str],
                   global_params: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """floodplanet.py:616-640 without the file loading / buffer padding: one sample in, the dict the
    DataLoader would collate out."""
    image, mean, std = normalize(image, norm_mode, global_params)
    img = apply_transforms(image, active, is_anno=False).float()
    tgt = apply_transforms(target, active, is_anno=True).long()
    return {"image": img, "target": tgt, "mean": mean, "std": std}
