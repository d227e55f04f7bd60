"""ORACLE (test infrastructure only) -- sliding-window inference restated on numpy.

Restates, for the inference configuration (BASELINE.json configs[4]):
  * the tile enumeration ``get_crop_slices(..., mode='exact')``
    (st_water_seg/datasets/utils.py:86-212, incl. the quirk at :203 where the bottom
    remainder rows use ``crop_height`` as the tile *width*),
  * the stitcher's scatter-add + weight canvas + ``/(w + 1e-5)`` + ``nan_to_num``
    (st_water_seg/utils/utils_image.py:410-494),
  * infer.py's post-processing: softmax over classes (:123), stitch (:160-163),
    ``np.clip(argmax, 0, 1) * 255`` as uint8 (:181-184).
Pinned against the reference itself: ``tests/golden/make_golden.py`` imports ``datasets/utils.py`` (tiler,
CropParams) and ``utils/utils_image.py`` (ImageStitcher_v2) by file path and stores their outputs
(``tests/golden/tiler.pt``, ``stitch.pt``); ``tests/test_oracle.py`` checks this module against them.
"""
from __future__ import annotations

from typing import List

import numpy as np


def crop_slices_exact(height: int, width: int, crop_height: int, crop_width: int,
                      step=None) -> List[List[int]]:
    """datasets/utils.py:86-212 with mode='exact': [h0, w0, h, w] per tile."""
    if step is None:
        h_step, w_step = crop_height, crop_width
    elif isinstance(step, tuple):
        h_step, w_step = step
    else:
        h_step = w_step = int(step)
    if h_step <= 0 or w_step <= 0:
        raise ValueError("step too small")
    if h_step > height or w_step > width:
        raise ValueError("step too large")
    n_h = 0
    while n_h * h_step + crop_height <= height:
        n_h += 1
    n_w = 0
    while n_w * w_step + crop_width <= width:
        n_w += 1
    out = [[i * h_step, j * w_step, crop_height, crop_width] for i in range(n_h) for j in range(n_w)]
    rem_h = height - n_h * h_step
    rem_w = width - n_w * w_step
    if rem_w != 0:
        out += [[i * h_step, n_w * w_step, crop_height, rem_w] for i in range(n_h)]
    if rem_h != 0:
        out += [[n_h * h_step, j * w_step, rem_h, crop_height] for j in range(n_w)]  # quirk :203
    if rem_h != 0 and rem_w != 0:
        out.append([n_h * h_step, n_w * w_step, rem_h, rem_w])
    return out


def softmax_np(x: np.ndarray, axis: int) -> np.ndarray:
    """scipy.special.softmax as called at infer.py:123."""
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


class Stitcher:
    """utils_image.py:410-494 for one scene: sum canvas + weight canvas."""

    def __init__(self, og_height: int, og_width: int, channels: int):
        self.canvas = np.zeros([og_height, og_width, channels], dtype=np.float32)
        self.weight = np.zeros([og_height, og_width], dtype=float)

    def add(self, pred_hwc: np.ndarray, h0: int, w0: int, height: int, width: int) -> None:
        hE = min(h0 + height, self.canvas.shape[0])
        wE = min(w0 + width, self.canvas.shape[1])
        dh, dw = hE - h0, wE - w0
        self.canvas[h0:hE, w0:wE, :] += pred_hwc[:dh, :dw, :]
        self.weight[h0:hE, w0:wE] += 1.0

    def combined(self) -> np.ndarray:
        return np.nan_to_num(self.canvas / (self.weight[:, :, None] + 1e-5))


def scene_mask_from_logits(tile_logits, tiles, og_height: int, og_width: int) -> np.ndarray:
    """infer.py:122-184 for one scene: per-tile logits [C,h,w] -> uint8 {0,255} mask."""
    c = tile_logits[0].shape[0]
    st = Stitcher(og_height, og_width, c)
    for lg, (h0, w0, hh, ww) in zip(tile_logits, tiles):
        pred = softmax_np(lg[None], axis=1)[0].transpose(1, 2, 0)
        st.add(pred, h0, w0, hh, ww)
    return (np.clip(st.combined().argmax(axis=2), 0, 1) * 255).astype('uint8')
