"""GPU parity tests of the fused normalise + augment kernel (csrc/augment.cu) through the C ABI:
against the golden outputs of the reference's own BaseDataset methods (tests/golden/augment.pt) and
against the oracle (oracle/augment_oracle.py) on seeded batches up to the BASELINE chip size.
Index work (which source pixel every output pixel reads, the int64 annotation) is bit-exact; the
fp32 image is bit-exact for norm_mode None / 'global' and within 2e-6 for 'local' (the per-plane
mean / std are fp64 reductions here, fp32 pairwise sums in numpy)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import augment_oracle as A

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def G():
    from floodplanet_code_b200 import augment
    return augment


def _active_from_golden(fx):
    return [{"transform": t["transform"], "anno": True, "kwargs": dict(t["kwargs"])} for t in fx["active"]]


def test_matches_reference_golden(G):
    for i, fx in enumerate(torch.load(GOLDEN / "augment.pt", weights_only=False)["cases"]):
        gp = None if fx["global_params"] is None else {k: v.numpy() for k, v in fx["global_params"].items()}
        aug = G.DeviceAugment(fx["cfg"], fx["norm_mode"], gp)
        batch = {"image": fx["image"][None].cuda(), "target": fx["target"].long()[None].cuda()}
        out = aug(batch, active=[_active_from_golden(fx)])
        assert torch.equal(out["target"][0].cpu(), fx["out_target"].long()), i
        if fx["norm_mode"] == "local":
            assert torch.allclose(out["image"][0].cpu(), fx["out_image"], rtol=2e-6, atol=2e-6), i
            # the index map itself is exact: zero-filled corners coincide
            assert torch.equal(out["image"][0].cpu() == 0, fx["out_image"] == 0), i
        else:
            assert torch.equal(out["image"][0].cpu(), fx["out_image"]), i
        assert torch.allclose(out["mean"].flatten().cpu(), fx["mean"], rtol=1e-6, atol=1e-7)
        assert torch.allclose(out["std"].flatten().cpu(), fx["std"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("shape", [(3, 4, 512, 512), (5, 6, 300, 300), (4, 4, 37, 53)])
def test_batch_matches_oracle(G, shape):
    n, c, h, w = shape
    rng = np.random.RandomState(n * 100 + h)
    image = rng.rand(n, c, h, w).astype(np.float32)
    target = (rng.rand(n, h, w) < 0.4).astype(np.int64) * 2 - (rng.rand(n, h, w) < 0.1)   # values -1..2
    cfg = {"hflip": {"active": True, "likelihood": 0.5}, "vflip": {"active": True, "likelihood": 0.5},
           "rotate": {"active": True, "likelihood": 0.7, "min_rot_angle": 0, "max_rot_angle": 360}}
    gp = {"mean": rng.rand(c), "std": rng.rand(c) + 0.5}
    np.random.seed(7)
    aug = G.DeviceAugment(cfg, "global", gp)
    active = aug.sample(n)
    out = aug({"image": torch.from_numpy(image).cuda(), "target": torch.from_numpy(target).cuda()}, active=active,
              c_pad=8 if c <= 8 else 16)
    for i in range(n):
        ref = A.augment_sample(image[i], target[i], active[i], "global", gp)
        assert torch.equal(out["target"][i].cpu(), ref["target"]), (i, active[i])
        assert torch.equal(out["image"][i].cpu(), ref["image"]), (i, active[i])
    # fused ingest output == bf16 cast of the fp32 output, channels beyond C zero
    nhwc = out["image_nhwc_bf16"].float().cpu()
    assert torch.equal(nhwc[..., :c], out["image"].cpu().permute(0, 2, 3, 1).to(torch.bfloat16).float())
    assert float(nhwc[..., c:].abs().max()) == 0.0 if nhwc.shape[-1] > c else True


def test_identity_flip_pairs_and_roundtrip(G):
    """Size-independent properties: no transform = identity; hflip twice = identity; rotate by 0 and by
    360 keep every pixel; 90-degree rotation four times returns the image."""
    g = torch.Generator(device="cuda").manual_seed(0)
    img = torch.rand(2, 4, 512, 512, generator=g, device="cuda")
    tgt = (torch.rand(2, 512, 512, generator=g, device="cuda") < 0.5).long()
    aug = G.DeviceAugment(None, None)
    out = aug({"image": img, "target": tgt})
    assert torch.equal(out["image"], img) and torch.equal(out["target"], tgt)
    hf = [{"transform": "hflip", "anno": True, "kwargs": {}}]
    once = aug({"image": img, "target": tgt}, active=[hf, hf])
    assert torch.equal(once["image"], img.flip(-1)) and torch.equal(once["target"], tgt.flip(-1))
    twice = aug({"image": once["image"], "target": once["target"]}, active=[hf, hf])
    assert torch.equal(twice["image"], img)
    for angle in (0.0, 360.0):
        r = [{"transform": "rotate", "anno": True, "kwargs": {"angle": angle}}]
        o = aug({"image": img, "target": tgt}, active=[r, r])
        assert torch.equal(o["image"], img) and torch.equal(o["target"], tgt)
    r90 = [{"transform": "rotate", "anno": True, "kwargs": {"angle": 90.0}}]
    cur = {"image": img, "target": tgt}
    for _ in range(4):
        cur = aug(cur, active=[r90, r90])
    assert torch.equal(cur["image"], img) and torch.equal(cur["target"], tgt)
    one = aug({"image": img, "target": tgt}, active=[r90, r90])
    assert torch.equal(one["image"], torch.rot90(img, 1, (-2, -1)))     # torchvision: counter-clockwise


def test_local_norm_statistics(G):
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.rand(3, 4, 300, 300, generator=g, device="cuda") * 5 + 2
    out = G.DeviceAugment(None, "local")({"image": img})
    flat = out["image"].reshape(3, 4, -1).double()
    assert float(flat.mean(-1).abs().max()) < 1e-5 and float((flat.std(-1, unbiased=False) - 1).abs().max()) < 1e-5
    ref_mean = img.reshape(3, 4, -1).double().mean(-1)
    assert torch.allclose(out["mean"].flatten(), ref_mean.flatten(), rtol=1e-6)


def test_errors(G):
    aug = G.DeviceAugment(None, None)
    with pytest.raises(RuntimeError):
        aug({"image": torch.rand(1, 4, 8, 8)})                       # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        aug({"image": torch.rand(1, 4, 8, 8, device="cuda"), "target": torch.zeros(1, 8, 8, device="cuda")})  # float target
    with pytest.raises(NotImplementedError):
        G.DeviceAugment(None, "bogus")
