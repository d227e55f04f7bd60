"""Tile-sharded sliding-window inference over one scene (BASELINE.json configs[4]).

What the reference does per scene (infer.py:64-184; predict.py uses ``cfg.crop_stride``):
enumerate crops with ``get_crop_slices(..., mode='exact')`` (datasets/utils.py:86-212), run the
model in eval mode on batches of crops, softmax the fp32 logits on the host, scatter-add them
into a canvas with a weight canvas (utils_image.py:410-494), divide by ``w + 1e-5``, and write
``np.clip(argmax, 0, 1) * 255`` as uint8.

Here the scene stays resident in HBM, crops are ingested straight from it (NCHW f32 ->
NHWC bf16, zero padded at the ragged edges), the UNet runs the folded-BatchNorm eval path, and
softmax / stitch / normalise / argmax / clip are two more kernels on the device.  Tiles are
independent, so N GPUs take contiguous ranges of the tile list with no data-path collective;
only the final uint8 masks (or, for overlapping strides, the canvases) are combined.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from .parallel import shard_range
from .unet import UNet


def crop_slices(height: int, width: int, crop_height: int, crop_width: int, step=None) -> List[List[int]]:
    """Tile list of the reference's 'exact' mode: full crops on the stride grid, then the right
    remainder column, the bottom remainder row, the corner -- each as [h0, w0, h, w].  Keeps the
    reference's quirk of using crop_height as the width of bottom-remainder tiles
    (datasets/utils.py:203); a tile never extends the scene, extents are clipped when stitched."""
    if step is None:
        hs, ws = crop_height, crop_width
    elif isinstance(step, tuple):
        hs, ws = int(step[0]), int(step[1])
    else:
        hs = ws = int(step)
    if hs <= 0 or ws <= 0:
        raise ValueError(f"Step of size {min(hs, ws)} is too small.")
    if hs > height or ws > width:
        raise ValueError(f"Step ({hs}, {ws}) is too large for image ({height}, {width})")
    nh = (height - crop_height) // hs + 1 if height >= crop_height else 0
    nw = (width - crop_width) // ws + 1 if width >= crop_width else 0
    tiles = [[i * hs, j * ws, crop_height, crop_width] for i in range(nh) for j in range(nw)]
    rem_h, rem_w = height - nh * hs, width - nw * ws
    if rem_w != 0:
        tiles += [[i * hs, nw * ws, crop_height, rem_w] for i in range(nh)]
    if rem_h != 0:
        tiles += [[nh * hs, j * ws, rem_h, crop_height] for j in range(nw)]
    if rem_h != 0 and rem_w != 0:
        tiles.append([nh * hs, nw * ws, rem_h, rem_w])
    return tiles


@torch.no_grad()
def predict_scene(unet: UNet, scene: torch.Tensor, crop: int = 512, stride: Optional[int] = None,
                  tile_batch: int = 32, rank: int = 0, world: int = 1, combine: bool = True):
    """Water mask (uint8 [H, W], values 0 / 255) of one scene [C, H, W] fp32.

    Each rank processes ``shard_range(n_tiles, rank, world)``.  With ``combine`` and an
    initialised process group the per-rank results are merged (canvases summed for overlapping
    strides, masks max-reduced otherwise); without it the caller gets this rank's partial mask.
    Returns (mask, n_tiles_processed, kernel_launches).
    """
    if not scene.is_cuda:
        raise RuntimeError("predict_scene: the scene must be resident on the CUDA device (no CPU fallback)")
    if unet.training:
        raise RuntimeError("predict_scene: call model.eval() first (reference: _set_model_to_eval)")
    c, H, W = scene.shape
    stride = crop if stride is None else stride
    tiles_all = crop_slices(H, W, crop, crop, stride)
    mine = [tiles_all[i] for i in shard_range(len(tiles_all), rank, world)]
    engine = unet._engine
    params = dict(unet.named_parameters())
    buffers = dict(unet.named_buffers())
    ncls = unet.n_classes
    dev = scene.device
    canvas = torch.zeros((H, W, ncls), dtype=torch.float32, device=dev)
    weight = torch.zeros((H, W), dtype=torch.float32, device=dev)
    scene = scene.contiguous()
    launches = 0
    for b0 in range(0, len(mine), tile_batch):
        chunk = mine[b0:b0 + tile_batch]
        # valid extents are clipped to the scene (the stitcher slices canvas[h0:hE, w0:wE])
        meta = [[h0, w0, min(hh, H - h0), min(ww, W - w0)] for h0, w0, hh, ww in chunk]
        tdev = torch.tensor(meta, dtype=torch.int32, device=dev)
        x = ops.ingest_scene_tiles(scene, tdev, crop, crop, engine.cin_pad)
        logits, _ = engine.forward(None, params, buffers, training=False, save=False, ingested=x)
        ops.softmax_stitch_add(logits, canvas, weight, tdev)
        launches += engine.launches + 1
    overlapping = stride < crop
    distributed = combine and world > 1 and torch.distributed.is_initialized()
    if distributed and overlapping:
        torch.distributed.all_reduce(canvas)
        torch.distributed.all_reduce(weight)
    mask = torch.empty((H, W), dtype=torch.uint8, device=dev)
    ops.canvas_to_mask_u8(canvas, weight, mask)
    launches += 1
    if distributed and not overlapping:
        torch.distributed.all_reduce(mask, op=torch.distributed.ReduceOp.MAX)
    return mask, len(mine), launches


_COPY_STREAMS = {}


def _row_len(tiles, t) -> int:
    """Number of tiles of `tiles` that start on the same scene row as tile t."""
    return sum(1 for u in tiles if u[0] == t[0])


def _copy_stream(dev) -> "torch.cuda.Stream":
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


@torch.no_grad()
def predict_scene_from_host(unet: UNet, scene_host: torch.Tensor, crop: int = 512, tile_batch: int = 32,
                            rank: int = 0, world: int = 1, mask_host: Optional[torch.Tensor] = None,
                            device: Optional[torch.device] = None):
    """End-to-end form of :func:`predict_scene` for a scene in HOST memory -- the shape of the
    reference loop (infer.py:112-184: `batch[key].to(device)` per batch :117-119, model, D2H :122,
    stitch :160-163, mask :181-184), non-overlapping tiles (infer.py:64-65: stride = crop).

    Each rank copies only the rows of the scene its contiguous tile range touches (host -> device, one
    tile batch at a time on a copy stream that runs one batch ahead of the compute stream; use pinned
    memory, and a `tile_batch` that is a multiple of the tiles per scene row so no row is copied twice),
    runs its tiles, stitches
    and thresholds on the device, and the uint8 band masks are max-reduced onto rank 0, which copies
    the full mask to `mask_host` (pinned uint8 [H, W], allocated if None).  No float logits, softmax
    or canvas ever cross PCIe.  Returns (mask_host or None on ranks > 0, n_tiles, launches,
    h2d_bytes, d2h_bytes)."""
    if scene_host.is_cuda:
        raise RuntimeError("predict_scene_from_host: the scene must be a host (ideally pinned) tensor")
    if unet.training:
        raise RuntimeError("predict_scene_from_host: call model.eval() first (reference: _set_model_to_eval)")
    dev = device if device is not None else next(unet.parameters()).device
    c, H, W = scene_host.shape
    tiles_all = crop_slices(H, W, crop, crop, crop)
    mine = [tiles_all[i] for i in shard_range(len(tiles_all), rank, world)]
    engine = unet._engine
    params = dict(unet.named_parameters())
    buffers = dict(unet.named_buffers())
    ncls = unet.n_classes
    mask = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    launches, h2d = 0, 0
    if mine:
        r0 = min(t[0] for t in mine)
        r1 = min(H, max(t[0] + t[2] for t in mine))
        canvas = torch.zeros((r1 - r0, W, ncls), dtype=torch.float32, device=dev)
        weight = torch.zeros((r1 - r0, W), dtype=torch.float32, device=dev)
        # host -> device copies run on their own stream, one tile batch ahead of the compute stream
        # (two row-band buffers, events both ways), so PCIe time hides behind the UNet of the previous batch
        # batches end on tile-row boundaries (tiles of one batch share as few scene rows as possible with the
        # next one): no scene row is copied twice, and the first -- un-hidden -- copy stays small
        batches, cur = [], []
        for t in mine:
            if cur and (len(cur) >= tile_batch or (t[0] != cur[-1][0] and len(cur) + _row_len(mine, t) > tile_batch)):
                batches.append(cur)
                cur = []
            cur.append(t)
        if cur:
            batches.append(cur)
        spans = [(min(t[0] for t in ch), min(H, max(t[0] + t[2] for t in ch))) for ch in batches]
        max_rows = max(e - s0 for s0, e in spans)
        bufs = [torch.empty((c, max_rows, W), dtype=torch.float32, device=dev) for _ in range(min(2, len(batches)))]
        compute = torch.cuda.current_stream(dev)
        copier = _copy_stream(dev)
        ready = [torch.cuda.Event() for _ in bufs]
        consumed = [torch.cuda.Event() for _ in bufs]
        for ev in consumed:
            ev.record(compute)

        def band_view(i):
            """Dense [C, rows_i, W] view of slot i % 2 (the ingest kernel addresses a contiguous scene)."""
            s0, e = spans[i]
            return bufs[i % len(bufs)].view(-1)[:c * (e - s0) * W].view(c, e - s0, W)

        def stage(i):
            s0, e = spans[i]
            slot = i % len(bufs)
            band = band_view(i)
            with torch.cuda.stream(copier):
                copier.wait_event(consumed[slot])
                for ch in range(c):                           # each plane slice is contiguous on the host
                    band[ch].copy_(scene_host[ch, s0:e], non_blocking=True)
                ready[slot].record(copier)
            return c * (e - s0) * W * 4

        h2d += stage(0)
        for i, chunk in enumerate(batches):
            if i + 1 < len(batches):
                h2d += stage(i + 1)
            slot = i % len(bufs)
            s0, e = spans[i]
            compute.wait_event(ready[slot])
            meta_in = [[h0 - s0, w0, min(hh, H - h0), min(ww, W - w0)] for h0, w0, hh, ww in chunk]
            x = ops.ingest_scene_tiles(band_view(i), torch.tensor(meta_in, dtype=torch.int32, device=dev),
                                       crop, crop, engine.cin_pad)
            consumed[slot].record(compute)                    # the band buffer may be overwritten from here on
            logits, _ = engine.forward(None, params, buffers, training=False, save=False, ingested=x)
            meta_out = [[h0 - r0, w0, min(hh, H - h0), min(ww, W - w0)] for h0, w0, hh, ww in chunk]
            ops.softmax_stitch_add(logits, canvas, weight, torch.tensor(meta_out, dtype=torch.int32, device=dev))
            launches += engine.launches + 2
        for buf in bufs:
            buf.record_stream(copier)
        # pixels of the band no tile of THIS rank covers keep weight 0 -> canvas 0 -> argmax 0 -> mask 0
        ops.canvas_to_mask_u8(canvas, weight, mask[r0:r1])
        launches += 1
    if world > 1 and torch.distributed.is_initialized():
        torch.distributed.reduce(mask, dst=0, op=torch.distributed.ReduceOp.MAX)
    d2h = 0
    out = None
    if rank == 0:
        out = mask_host if mask_host is not None else torch.empty((H, W), dtype=torch.uint8, pin_memory=True)
        out.copy_(mask, non_blocking=False)                   # device -> host, synchronises
        d2h = mask.numel()
    return out, len(mine), launches, h2d, d2h
