"""Edge cases at the module boundary (GPU): minimum size, memory formats, dtypes the reference rejects,
frozen-BatchNorm backward through the late-fusion and encode/decode seams."""
import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def _unet(c=4, seed=0):
    from floodplanet_code_b200.unet import UNet
    m = UNet(c, 3)
    m.load_state_dict(O.init_state_dict(c, 3, seed=seed))
    return m.cuda()


def test_minimum_size_16x16_and_memory_formats():
    from floodplanet_code_b200.loss import MaskedCrossEntropyLoss
    m = _unet().train()
    b = O.synthetic_batch(4, 4, 16, 16, seed=3, block=4, device="cuda")
    logits = m(b["image"])
    assert logits.shape == (4, 3, 16, 16) and torch.isfinite(logits).all()
    MaskedCrossEntropyLoss(0)(logits, b["target"]).backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())
    with pytest.raises(RuntimeError, match="at least 16x16"):
        m(torch.rand(1, 4, 15, 20, device="cuda"))
    # channels_last, sliced (non-contiguous) and float64 inputs give the same logits as the contiguous fp32 tensor
    m.eval()
    x = torch.rand(2, 4, 48, 40, device="cuda")
    with torch.no_grad():
        want = m(x)
        assert torch.equal(m(x.contiguous(memory_format=torch.channels_last)), want)
        wide = torch.rand(2, 6, 48, 40, device="cuda")
        wide[:, 1:5] = x
        assert torch.equal(m(wide[:, 1:5]), want)
        assert torch.equal(m(x.double()), want)          # cast to fp32 on ingest, like `.float()` in the dataset


def test_targets_must_be_int64_like_the_reference():
    from floodplanet_code_b200.loss import MaskedCrossEntropyLoss
    logits = torch.randn(1, 3, 16, 16, device="cuda")
    with pytest.raises(RuntimeError):
        MaskedCrossEntropyLoss(0)(logits, torch.zeros(1, 16, 16, dtype=torch.int32, device="cuda"))
    # out-of-range targets are counted (and raise on request), never silently used
    tgt = torch.full((1, 16, 16), 7, dtype=torch.int64, device="cuda")
    with pytest.raises(IndexError):
        MaskedCrossEntropyLoss(0, check_targets=True)(logits, tgt)


def test_frozen_batchnorm_backward_through_late_fusion_and_seams():
    """eval() + grad mode: LateFusionModel and UNet.encode / decode differentiate with running statistics; BatchNorm
    buffers stay untouched; gradients agree with the fp32 oracle's eval-mode autograd at the layers next to the loss."""
    from floodplanet_code_b200.lf_model import LateFusionModel
    torch.manual_seed(0)
    lf = LateFusionModel({"ms_image": 4, "dem": 1}, 3, 1e-3, ignore_index=0).cuda()
    b = O.synthetic_batch(2, 4, 32, 32, seed=5, block=8, device="cuda")
    b["dem"] = torch.rand(2, 1, 32, 32, device="cuda")
    lf._set_model_to_train()
    with torch.no_grad():
        for _ in range(2):
            lf(b)
    lf._set_model_to_eval()
    before = {k: v.clone() for k, v in lf.state_dict().items()}
    out = lf(b)
    assert out.requires_grad
    loss = lf.loss_func(out, b["target"])
    loss.backward()
    for k, v in lf.state_dict().items():
        if "running" in k or "num_batches" in k:
            assert torch.equal(v, before[k]), k
    sd = {k: v.detach().clone() for k, v in lf.state_dict().items()}
    keys = O.trainable_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
    ologits = O.lf_forward(sd, b, training=False)
    oloss, _ = O.masked_ce(ologits, b["target"], 0)
    oloss.backward()
    assert abs(float(loss.detach()) - float(oloss.detach())) <= 1e-2 * abs(float(oloss.detach()))
    named = dict(lf.named_parameters())
    for k in ("decoder.outc.conv.weight", "decoder.outc.conv.bias", "decoder.up4.conv.double_conv.4.weight",
              "decoder.up4.conv.double_conv.3.bias"):
        g, og = named[k].grad.double().flatten(), sd[k].grad.double().flatten()
        assert float((g - og).norm() / og.norm()) < 2e-2, k
    # conv biases get REAL gradients in frozen mode (they are exactly zero in training mode)
    assert float(named["encoders.ms_image.inc.double_conv.0.bias"].grad.abs().max()) > 0
    # encode / decode seam
    m = _unet(seed=2).eval()
    x = torch.rand(1, 4, 32, 32, device="cuda")
    feats = m.encode(x)
    assert all(f.requires_grad for f in feats)
    m.decode(feats).sum().backward()
    assert float(m.inc.double_conv[0].weight.grad.abs().max()) > 0
