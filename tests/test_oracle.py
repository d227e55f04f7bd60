"""CPU tests (no GPU): the oracle restatement against the golden vectors generated from the
reference itself (tests/golden/make_golden.py), plus the op-semantics edge cases of
SURVEY.md section 8(c).  These pin the oracle; the GPU parity tests then compare the CUDA
path with the oracle and with the same fixtures."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import tiling_oracle as T
from oracle import unet_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ["unet_c4_32", "unet_c4_44x36", "unet_c6_37_ef", "unet_c4_32_allignored"]


def load(name):
    return torch.load(GOLDEN / f"{name}.pt", weights_only=False)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    fx = load(name)
    cfg = fx["cfg"]
    sd = O.init_state_dict(cfg["c"], cfg["n_classes"], seed=cfg["seed"])
    # identical initialisation (same RNG consumption order as the reference constructor)
    for k, v in fx["init_checksum"].items():
        assert float(sd[k].double().sum()) == pytest.approx(v, rel=1e-12, abs=1e-12), k
    assert list(sd.keys()) == list(fx["init_checksum"].keys())  # state_dict key order too
    batch = {"image": fx["image"], "target": fx["target"]}
    loss, pred, logits, grads = O.training_step(sd, batch, cfg["ignore_index"], early_fusion=False)
    assert torch.allclose(logits, fx["logits_train"], rtol=1e-5, atol=1e-6)
    assert float(loss) == pytest.approx(fx["loss"], rel=1e-5, abs=1e-7)
    assert torch.equal(pred, fx["pred"])
    for k, g in fx["grads"].items():
        assert float(grads[k].double().norm()) == pytest.approx(g["norm"], rel=2e-4, abs=1e-9), k
    for k, g in fx["grad_full"].items():
        assert torch.allclose(grads[k], g, rtol=1e-3, atol=1e-7), k
    for k, v in fx["bn_after"].items():
        assert torch.allclose(sd[k].to(v.dtype), v, rtol=1e-5, atol=1e-7), k
    with torch.no_grad():
        ev = O.unet_forward(sd, fx["image"], training=False)
    assert torch.allclose(ev, fx["logits_eval"], rtol=1e-4, atol=1e-5)


def test_all_ignored_batch_gives_zero_loss_and_zero_grads():
    fx = load("unet_c4_32_allignored")
    assert fx["loss"] == 0.0
    assert all(g["norm"] == 0.0 for g in fx["grads"].values())


def test_op_semantics_golden():
    fx = load("op_semantics")
    assert fx["maxpool_zero_idx"] == [0, 2, 8, 10]          # first element of each window
    assert fx["maxpool_nan_idx"][0] == 5 and fx["maxpool_nan_isnan"][0]  # NaN wins
    assert fx["argmax_ties"] == [1, 0, 0, 0]                 # NaN maximal, ties -> lowest index
    assert fx["ce_all_ignored_isnan"] and fx["ce_all_ignored_grad_abs_sum"] == 0.0
    assert torch.allclose(fx["bilinear_4_to_8"], torch.tensor([i * 3 / 7 for i in range(8)]), atol=1e-6)
    x = fx["bn_x"]
    mean = x.mean((0, 2, 3))
    var_unbiased = x.var((0, 2, 3), unbiased=True)
    assert torch.allclose(fx["bn_running_mean"], 0.1 * mean, atol=1e-6)
    assert torch.allclose(fx["bn_running_var"], 0.9 + 0.1 * var_unbiased, atol=1e-6)


def test_early_fusion_order():
    b = {"image": torch.zeros(1, 4, 2, 2), "hand": torch.full((1, 1, 2, 2), 5.0),
         "dem": torch.full((1, 1, 2, 2), 1.0), "slope": torch.full((1, 1, 2, 2), 2.0)}
    x = O.early_fusion_input(b)
    assert x.shape[1] == 7 and x[0, 4:, 0, 0].tolist() == [1.0, 2.0, 5.0]


def test_tiler_matches_reference_golden():
    fx = load("tiler")
    for (H, W, ch, cw, st), want in fx.items():
        sl = T.crop_slices_exact(H, W, ch, cw, st)
        assert len(sl) == want["count"]
        assert sl[:5] == want["head"] and sl[-5:] == want["tail"]
        chk = int(sum((i + 1) * (a + 3 * b + 5 * c + 7 * d) for i, (a, b, c, d) in enumerate(sl)))
        assert chk == want["checksum"]
    assert len(T.crop_slices_exact(10240, 10240, 512, 512, 512)) == 400


def test_stitch_nonoverlapping_equals_argmax_of_logits():
    rng = np.random.default_rng(0)
    tiles = T.crop_slices_exact(96, 64, 32, 32, 32)
    logits = [rng.standard_normal((3, 32, 32)).astype(np.float32) for _ in tiles]
    mask = T.scene_mask_from_logits(logits, tiles, 96, 64)
    for lg, (h0, w0, hh, ww) in zip(logits, tiles):
        want = (np.clip(lg.argmax(0), 0, 1) * 255).astype("uint8")
        assert np.array_equal(mask[h0:h0 + hh, w0:w0 + ww], want)


def test_micro_metrics_formulas():
    conf = torch.tensor([[0, 0, 0], [3, 5, 2], [0, 0, 0]])
    m = O.micro_metrics(conf)
    assert m["Accuracy"] == pytest.approx(0.5) and m["F1"] == pytest.approx(0.5)
    assert m["Jaccard"] == pytest.approx(5 / 15)
    # torchmetrics >= 0.11 micro Jaccard with ignore_index inside [0, C): valid pixels PREDICTED as the
    # ignored class leave the denominator (hand-worked: tp 5, total 10, 3 predicted as class 0)
    m0 = O.micro_metrics(conf, ignore_index=0)
    assert m0["Jaccard"] == pytest.approx(5 / (2 * 10 - 5 - 3))
    assert m0["Accuracy"] == pytest.approx(0.5)
    assert O.micro_metrics(conf, ignore_index=-100)["Jaccard"] == pytest.approx(5 / 15)


def test_late_fusion_oracle_matches_reference_golden():
    """Fixture made by the reference's own LateFusionModel (lf_model.py) on CPU fp32."""
    fx = load("lf_c4_dem1_32")
    cfg = fx["cfg"]
    sd = O.init_lf_state_dict(cfg["in_channels"], cfg["n_classes"], seed=cfg["seed"])
    assert list(sd.keys()) == fx["state_dict_keys"]          # state_dict key order of the reference
    for k, v in fx["init_checksum"].items():
        assert float(sd[k].double().sum()) == pytest.approx(v, rel=1e-12, abs=1e-12), k
    loss, pred, logits, grads = O.lf_training_step(sd, fx["batch"], cfg["ignore_index"])
    assert torch.allclose(logits, fx["logits_train"], rtol=1e-5, atol=1e-6)
    assert float(loss) == pytest.approx(fx["loss"], rel=1e-5, abs=1e-7)
    assert torch.equal(pred, fx["pred"])
    for k, g in fx["grads"].items():
        assert float(grads[k].double().norm()) == pytest.approx(g["norm"], rel=2e-4, abs=1e-9), k
    for k, g in fx["grad_full"].items():
        assert torch.allclose(grads[k], g, rtol=1e-3, atol=1e-7), k
    for k, v in fx["bn_after"].items():
        assert torch.allclose(sd[k].to(v.dtype), v, rtol=1e-5, atol=1e-7), k
    with torch.no_grad():
        ev = O.lf_forward(sd, fx["batch"], training=False)
        feats = O.unet_encode(sd, fx["batch"]["image"], training=False, prefix="encoders.ms_image.")
    assert torch.allclose(ev, fx["logits_eval"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(feats[4], fx["enc_feat_eval_x5"], rtol=1e-4, atol=1e-5)
    for f, want in zip(feats, fx["enc_feat_eval_checksum"]):
        assert float(f.double().sum()) == pytest.approx(want, rel=1e-4)


def test_encode_decode_compose_to_forward():
    sd = O.init_state_dict(4, 3, seed=2)
    x = torch.rand(1, 4, 32, 32, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        a = O.unet_forward(sd, x, training=False)
        b = O.unet_decode(sd, O.unet_encode(sd, x, training=False), training=False)
    assert torch.equal(a, b)


def test_stitch_oracle_matches_reference_image_stitcher_golden():
    """oracle/tiling_oracle.py (tiler + softmax + Stitcher + mask rule) against outputs of the reference's
    OWN ImageStitcher_v2 / get_crop_slices / CropParams and scipy softmax (tests/golden/stitch.pt, made by
    make_golden.py stitch): canvas to 1e-12, uint8 mask bit-exact -- stride = crop, ragged and overlapping."""
    fx = load("stitch")
    for cs in fx["cases"]:
        H, W, crop, stride, ncls = cs["H"], cs["W"], cs["crop"], cs["stride"], cs["n_classes"]
        tiles = T.crop_slices_exact(H, W, crop, crop, stride)
        assert tiles == cs["tiles"]
        rng = np.random.RandomState(cs["seed"])
        logits = [rng.standard_normal((ncls, crop, crop)).astype(np.float32) * 2 for _ in tiles]
        assert float(np.stack(logits).astype(np.float64).sum()) == cs["logits_checksum"]
        st = T.Stitcher(H, W, ncls)
        for lg, (h0, w0, hh, ww) in zip(logits, tiles):
            st.add(T.softmax_np(lg[None], axis=1)[0].transpose(1, 2, 0), h0, w0, hh, ww)
        canvas = st.combined()
        want = cs["canvas"].numpy()
        assert canvas.dtype == want.dtype and np.array_equal(canvas, want)     # bit-exact, scipy softmax included
        mask = T.scene_mask_from_logits(logits, tiles, H, W)
        assert np.array_equal(mask, cs["mask"].numpy())
