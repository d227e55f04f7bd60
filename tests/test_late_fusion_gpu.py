"""Late fusion and the encode/decode API (GPU) against the oracle and the fixture written by the
reference's own ``lf_model.LateFusionModel`` (tests/golden/make_golden.py lf).

Same bars as tests/test_unet_gpu.py: loss within 1e-2 of the fp32 reference; whole-network
logits inside the bf16-storage envelope and no further from the fp32 reference than the
bf16-emulated oracle; the fusion-conv and head gradients (next to the loss / identical inputs)
within 2e-2."""
from pathlib import Path

import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"

LOGIT_TOL = 1e-2
GRAD_TOL = 2e-2
# Whole-network envelope (see tests/test_unet_gpu.py).  Late fusion adds a second encoder and
# five fusion convs in front of the decoder, so at random init the bf16-STORAGE distance to the
# fp32 reference is larger than for the plain UNet: the fp32-arithmetic oracle with only the
# storage points emulated sits at 9.0 % on this fixture (measured), the CUDA path at 9.1 %, and
# the two are 3.5 % apart.  The assertion that matters is the relative one below.
NET_LOGIT_ENVELOPE = 12e-2
NET_VS_EMULATED = 4e-2


def rel(a, b):
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_late_fusion_train_step_matches_reference_golden():
    from floodplanet_code_b200.water_seg_model import build_model
    fx = torch.load(GOLDEN / "lf_c4_dem1_32.pt", weights_only=False)
    cfg = fx["cfg"]
    sd = O.init_lf_state_dict(cfg["in_channels"], cfg["n_classes"], seed=cfg["seed"])
    model = build_model("lf_model", dict(cfg["in_channels"]), cfg["n_classes"], 1e-4, 50, None,
                        cfg["ignore_index"])
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    batch = {k: v.cuda() for k, v in fx["batch"].items()}
    model._set_model_to_train()
    logits = model.forward(batch)
    assert logits.dtype == torch.float32 and logits.shape == fx["logits_train"].shape
    sd_emu = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        emu = O.lf_forward_bf16_emulated(sd_emu, fx["batch"], True)
    e_fp32, e_emu, inherent = rel(logits, fx["logits_train"]), rel(logits, emu), rel(emu, fx["logits_train"])
    print(f"late fusion: logits vs fp32 ref {e_fp32:.4f}, vs bf16-emulated {e_emu:.4f}, inherent {inherent:.4f}")
    assert e_fp32 < NET_LOGIT_ENVELOPE
    assert e_fp32 < 1.5 * inherent + LOGIT_TOL
    assert e_emu < NET_VS_EMULATED
    loss = model.loss_func(logits, batch["target"])
    assert abs(float(loss) - fx["loss"]) <= LOGIT_TOL * abs(fx["loss"])
    # argmax differs from the fp32 reference only where logits nearly tie: no more often than
    # for the bf16-emulated fp32-arithmetic oracle (95.2 % agreement on this fixture)
    agree = float((model.loss_func.last_pred.cpu() == fx["pred"]).float().mean())
    agree_emu = float((emu.argmax(1) == fx["pred"]).float().mean())
    assert agree > agree_emu - 0.02 and agree > 0.93, (agree, agree_emu)
    loss.backward()
    _, _, _, ograds = O.lf_training_step(sd, fx["batch"], cfg["ignore_index"])
    named = dict(model.named_parameters())
    for k, g in ograds.items():
        got = named[k].grad
        assert got is not None and got.dtype == torch.float32 and got.shape == g.shape, k
        if k.endswith(".0.bias") or k.endswith(".3.bias"):      # conv bias feeding a train-mode BN
            continue
        a, b = got.detach().double().flatten().cpu(), g.double().flatten()
        cos = float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))
        ratio = float(a.norm() / b.norm().clamp_min(1e-30))
        assert cos > 0.75 and 0.8 < ratio < 1.25, (k, cos, ratio)
    for k in ("decoder.outc.conv.weight", "decoder.outc.conv.bias", "decoder.up4.conv.double_conv.4.weight"):
        assert rel(named[k].grad, ograds[k]) < GRAD_TOL, k
    # BN buffers of the dem encoder's first layer (un-amplified inputs)
    after = model.state_dict()
    for k, v in fx["bn_after"].items():
        if k.endswith("num_batches_tracked"):
            assert int(after[k]) == int(v)
        else:
            assert rel(after[k], v) < 1e-2, k
    # eval mode (folded BatchNorm) against the reference's eval logits
    model._set_model_to_eval()
    with torch.no_grad():
        ev = model.forward(batch)
    assert rel(ev, fx["logits_eval"]) < NET_LOGIT_ENVELOPE


def test_fusion_conv_gradients_at_identical_inputs():
    """The 1x1 fusion stage in isolation through the engine's kernels: identical (bf16) inputs,
    fp32 torch reference -> north_star bars (1e-2 forward, 2e-2 gradients)."""
    from floodplanet_code_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    n, h, w, fs, k = 2, 24, 20, 128, 3
    feats = (torch.randn(n, h, w, fs * k, generator=g, device="cuda")).relu().to(torch.bfloat16)
    wt = (torch.randn(fs, fs * k, 1, 1, generator=g, device="cuda") / (fs * k) ** 0.5).to(torch.bfloat16).float()
    bias = torch.randn(fs, generator=g, device="cuda")
    dy = torch.randn(n, h, w, fs, generator=g, device="cuda").to(torch.bfloat16)
    y = torch.empty(n, h, w, fs, dtype=torch.bfloat16, device="cuda")
    ops.conv1x1(feats, ops.repack_1x1(wt, False), y, torch.ones(fs, device="cuda"), bias)
    x32 = feats.float().permute(0, 3, 1, 2).requires_grad_(True)
    w32 = wt.clone().requires_grad_(True)
    b32 = bias.clone().requires_grad_(True)
    ref = torch.nn.functional.conv2d(x32, w32, b32)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    assert rel(y.float().permute(0, 3, 1, 2), ref) < LOGIT_TOL
    dx = torch.empty_like(feats)
    ops.conv1x1(dy, ops.repack_1x1(wt, True), dx)
    assert rel(dx.float().permute(0, 3, 1, 2), x32.grad) < GRAD_TOL
    dw = torch.empty_like(wt)
    ws = torch.empty(ops.conv1x1_wgrad_workspace_bytes(n, h, w, fs * k, fs) // 4, device="cuda")
    ops.conv1x1_wgrad(feats, dy, dw, ws)
    assert rel(dw, w32.grad) < GRAD_TOL
    db = torch.empty(fs, device="cuda")
    ops.channel_sum(dy, db)
    assert rel(db, b32.grad) < GRAD_TOL


@pytest.mark.parametrize("hw", [(32, 32), (44, 36)])
def test_encode_decode_api_matches_forward(hw):
    """UNet.encode / UNet.decode (unet.py:113-131) reproduce UNet.forward -- bit-exact in eval
    mode (the features cross the seam as bf16-representable fp32) -- and train through autograd."""
    from floodplanet_code_b200.unet import UNet
    h, w = hw
    sd = O.init_state_dict(4, 3, seed=4)
    net = UNet(4, 3)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    x = torch.rand(2, 4, h, w, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        full = net(x)
        feats = net.encode(x)
        assert [tuple(f.shape[1:]) for f in feats] == [(64, h, w), (128, h // 2, w // 2), (256, h // 4, w // 4),
                                                       (512, h // 8, w // 8), (512, h // 16, w // 16)]
        assert all(f.dtype == torch.float32 for f in feats)
        two = net.decode(feats)
    assert torch.equal(full, two)
    with torch.no_grad():
        ofeats = O.unet_encode(sd, x.cpu(), training=False)
    assert rel(feats[0], ofeats[0]) < LOGIT_TOL          # first block: un-amplified inputs
    # training through the two halves gives the same gradients as the fused network
    net.train()
    t = (torch.rand(2, h, w) > 0.5).long().cuda()
    from floodplanet_code_b200.loss import MaskedCrossEntropyLoss
    loss_fn = MaskedCrossEntropyLoss(ignore_index=None)
    net.zero_grad()
    loss_fn(net(x), t).backward()
    g_full = {k: p.grad.clone() for k, p in net.named_parameters()}
    net.load_state_dict(sd, strict=True)                  # reset BN buffers
    net.zero_grad()
    loss_fn(net.decode(net.encode(x)), t).backward()
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        if k.endswith(".0.bias") or k.endswith(".3.bias"):
            continue
        # identical kernels on identical data except one extra bf16 rounding of the feature
        # gradients at the seam
        assert rel(p.grad, g_full[k]) < GRAD_TOL, k


def test_encoder_decoder_modules_standalone():
    from floodplanet_code_b200.unet import UNetDecoder, UNetEncoder
    torch.manual_seed(0)
    enc, dec = UNetEncoder(5).cuda().eval(), UNetDecoder(2).cuda().eval()
    x = torch.rand(1, 5, 48, 48, device="cuda")
    with torch.no_grad():
        feats = enc(x)
        logits = dec(feats)
        out = dec.get_output_feats(feats)
    assert logits.shape == (1, 2, 48, 48) and out.shape == (1, 64, 48, 48)
    sd = {**{k: v.cpu() for k, v in enc.state_dict().items()}, **{k: v.cpu() for k, v in dec.state_dict().items()}}
    with torch.no_grad():
        ofe = O.unet_encode(sd, x.cpu(), training=False)
        # decoder parity at IDENTICAL inputs: feed the oracle the features the CUDA encoder produced
        ologits = O.unet_decode(sd, [f.cpu() for f in feats], training=False)
        oout = O.unet_decode(sd, [f.cpu() for f in feats], training=False, head=False)
    assert rel(feats[0], ofe[0]) < LOGIT_TOL
    assert rel(logits, ologits) < NET_VS_EMULATED
    assert rel(out, oout) < NET_VS_EMULATED
