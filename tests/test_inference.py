"""Sliding-window inference: tiler vs the reference golden (CPU) and the on-device
softmax/stitch/argmax/clip post-processing + tile sharding vs the numpy oracle (GPU)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import tiling_oracle as T
from oracle import unet_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_product_tiler_matches_reference_golden():
    from floodplanet_code_b200.inference import crop_slices
    fx = torch.load(GOLDEN / "tiler.pt", weights_only=False)
    for (H, W, ch, cw, st), want in fx.items():
        sl = crop_slices(H, W, ch, cw, st)
        assert len(sl) == want["count"] and sl[:5] == want["head"] and sl[-5:] == want["tail"]
        assert int(sum((i + 1) * (a + 3 * b + 5 * c + 7 * d) for i, (a, b, c, d) in enumerate(sl))) == want["checksum"]
    with pytest.raises(ValueError):
        crop_slices(100, 100, 64, 64, 0)


def _model():
    from floodplanet_code_b200.unet import UNet
    m = UNet(4, 3)
    m.load_state_dict(O.init_state_dict(4, 3, seed=0))
    m = m.cuda()
    m.train()
    with torch.no_grad():
        m(O.synthetic_batch(2, 4, 64, 64, seed=3, device="cuda")["image"])  # non-trivial running stats
    return m.eval()


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,crop,stride", [(96, 64, 32, 32), (80, 112, 32, 16), (100, 70, 32, 32)])
def test_predict_scene_matches_oracle_stitch(H, W, crop, stride):
    from floodplanet_code_b200.inference import crop_slices, predict_scene
    m = _model()
    g = torch.Generator().manual_seed(5)
    scene = torch.rand(4, H, W, generator=g).cuda()
    mask, n_tiles, launches = predict_scene(m, scene, crop=crop, stride=stride, tile_batch=5)
    tiles = crop_slices(H, W, crop, crop, stride)
    assert n_tiles == len(tiles) and launches > 0
    # oracle post-processing on the SAME per-tile logits (from the module's public forward on
    # zero-padded crops, exactly what the reference's dataset + infer loop feed the model)
    tile_logits = []
    for h0, w0, hh, ww in tiles:
        crop_img = torch.zeros(1, 4, crop, crop, device="cuda")
        vh, vw = min(hh, H - h0, crop), min(ww, W - w0, crop)
        crop_img[0, :, :vh, :vw] = scene[:, h0:h0 + vh, w0:w0 + vw]
        with torch.no_grad():
            tile_logits.append(m(crop_img)[0].cpu().numpy())
    want = T.scene_mask_from_logits(tile_logits, tiles, H, W)
    got = mask.cpu().numpy()
    if stride == crop:
        assert np.array_equal(got, want)           # bit-exact water mask
    else:
        assert (got != want).mean() < 2e-3         # fp32 summation order of overlapping tiles


@pytest.mark.gpu
def test_tile_sharding_is_a_partition_of_the_single_gpu_result():
    from floodplanet_code_b200.inference import predict_scene
    m = _model()
    scene = torch.rand(4, 128, 96, generator=torch.Generator().manual_seed(6)).cuda()
    full, n, _ = predict_scene(m, scene, crop=32, stride=32)
    parts = [predict_scene(m, scene, crop=32, stride=32, rank=r, world=3, combine=False) for r in range(3)]
    assert sum(p[1] for p in parts) == n == 12
    merged = torch.stack([p[0] for p in parts]).max(0).values
    assert torch.equal(merged, full)


@pytest.mark.gpu
def test_device_stitch_kernels_match_reference_stitcher_golden():
    """softmax_stitch_add + canvas_to_mask_u8 on the logits of tests/golden/stitch.pt against the outputs
    of the reference's OWN ImageStitcher_v2 + infer.py mask rule stored there."""
    from floodplanet_code_b200 import ops
    fx = torch.load(GOLDEN / "stitch.pt", weights_only=False)
    for cs in fx["cases"]:
        H, W, crop, stride, ncls = cs["H"], cs["W"], cs["crop"], cs["stride"], cs["n_classes"]
        tiles = cs["tiles"]
        rng = np.random.RandomState(cs["seed"])
        logits = np.stack([rng.standard_normal((ncls, crop, crop)).astype(np.float32) * 2 for _ in tiles])
        assert float(logits.astype(np.float64).sum()) == cs["logits_checksum"]
        meta = [[h0, w0, min(hh, H - h0), min(ww, W - w0)] for h0, w0, hh, ww in tiles]
        tdev = torch.tensor(meta, dtype=torch.int32, device="cuda")
        canvas = torch.zeros((H, W, ncls), dtype=torch.float32, device="cuda")
        weight = torch.zeros((H, W), dtype=torch.float32, device="cuda")
        ops.softmax_stitch_add(torch.from_numpy(logits).cuda(), canvas, weight, tdev)
        mask = torch.empty((H, W), dtype=torch.uint8, device="cuda")
        ops.canvas_to_mask_u8(canvas, weight, mask)
        want_canvas = cs["canvas"].numpy()
        got_canvas = (canvas.double() / (weight.double()[:, :, None] + 1e-5)).cpu().numpy()
        assert np.abs(got_canvas - want_canvas).max() < 2e-6          # fp32 exp / summation order
        got, want = mask.cpu().numpy(), cs["mask"].numpy()
        if stride == crop:
            assert np.array_equal(got, want)
        else:
            assert (got != want).mean() < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,crop", [(96, 64, 32), (100, 70, 32)])
def test_predict_scene_from_host_equals_device_resident_path(H, W, crop):
    """The end-to-end form (scene in pinned host memory, per-rank row band H2D, uint8 mask D2H; the shape of
    infer.py:112-184) gives bit-identical masks to the device-resident path, also when tile-sharded."""
    from floodplanet_code_b200.inference import predict_scene, predict_scene_from_host
    m = _model()
    scene = torch.rand(4, H, W, generator=torch.Generator().manual_seed(8))
    want, n, _ = predict_scene(m, scene.cuda(), crop=crop, stride=crop, tile_batch=5)
    host = scene.pin_memory()
    got, n1, launches, h2d, d2h = predict_scene_from_host(m, host, crop=crop, tile_batch=5)
    # (a batch that starts inside a tile row re-copies that row band: >=; equality for aligned batches below)
    assert n1 == n and launches > 0 and h2d >= scene.numel() * 4 and d2h == H * W
    assert got.dtype == torch.uint8 and not got.is_cuda and torch.equal(got, want.cpu())
    # batches aligned to whole tile rows copy every scene row exactly once, and several batches (two band
    # buffers, copy stream one batch ahead) give the same mask as one
    if H % crop == 0 and W % crop == 0:
        per_row = W // crop
        got2, _, _, h2d2, _ = predict_scene_from_host(m, host, crop=crop, tile_batch=per_row)
        assert h2d2 == scene.numel() * 4 and torch.equal(got2, got)
    # each rank of a 3-way shard copies only the rows its tiles touch
    from floodplanet_code_b200.inference import crop_slices
    from floodplanet_code_b200.parallel import shard_range
    tiles_all = crop_slices(H, W, crop, crop, crop)
    for r in range(3):
        mine = [tiles_all[i] for i in shard_range(len(tiles_all), r, 3)]
        rows = min(H, max(t[0] + t[2] for t in mine)) - min(t[0] for t in mine)
        _, k, _, hb, _ = predict_scene_from_host(m, host, crop=crop, tile_batch=64, rank=r, world=3)
        assert k == len(mine) and hb == 4 * rows * W * 4


@pytest.mark.gpu
def test_infer_py_shaped_loop_through_the_model_registry():
    """The caller's side of the seam, written the way the reference's infer.py:86-184 drives it: build_model
    by registry name, `_set_model_to_eval()`, `.to('cuda')`, per batch `batch[key].to(device)`,
    `model(batch).detach().cpu().numpy()` (fp32 logits for numpy), scipy softmax, `b c h w -> b h w c`,
    ImageStitcher-style accumulation (the pinned oracle Stitcher), `np.clip(argmax, 0, 1) * 255`.  The mask
    must equal the tile-sharded device path's."""
    from scipy.special import softmax
    from floodplanet_code_b200.inference import crop_slices, predict_scene
    from floodplanet_code_b200.water_seg_model import build_model
    H, W, crop = 96, 128, 32
    model = build_model("ef_model", {"ms_image": 4, "dem": 1}, 3, 1e-4, 50, None, 0)
    model.model.load_state_dict(O.init_state_dict(5, 3, seed=2))
    model._set_model_to_eval()
    model = model.to("cuda")
    device = "cuda"
    g = torch.Generator().manual_seed(11)
    scene = torch.rand(5, H, W, generator=g)
    tiles = crop_slices(H, W, crop, crop, crop)
    st = T.Stitcher(H, W, 3)
    with torch.no_grad():
        for b0 in range(0, len(tiles), 5):                       # DataLoader batches of crops + metadata
            chunk = tiles[b0:b0 + 5]
            batch = {"image": torch.stack([scene[:4, h0:h0 + hh, w0:w0 + ww] for h0, w0, hh, ww in chunk]),
                     "dem": torch.stack([scene[4:5, h0:h0 + hh, w0:w0 + ww] for h0, w0, hh, ww in chunk]),
                     "metadata": [{"crop_params": t} for t in chunk]}
            for key, value in batch.items():
                if isinstance(value, torch.Tensor):
                    batch[key] = value.to(device)
            output = model(batch).detach().cpu().numpy()
            assert output.dtype == np.float32 and output.shape == (len(chunk), 3, crop, crop)
            preds = softmax(output, axis=1).transpose(0, 2, 3, 1)
            for b, (h0, w0, hh, ww) in enumerate(chunk):
                st.add(preds[b], h0, w0, hh, ww)
    want = (np.clip(st.combined().argmax(axis=2), 0, 1) * 255).astype("uint8")
    # the device path: same UNet object, early-fusion channels pre-concatenated in the scene
    got, n, _ = predict_scene(model.model, scene.cuda(), crop=crop, stride=crop, tile_batch=7)
    assert n == len(tiles) and np.array_equal(got.cpu().numpy(), want)
