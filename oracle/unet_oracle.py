"""ORACLE -- test infrastructure only, never a product path.

A CPU/fp32 restatement of the reference algorithm for the UNet hot path, written functionally
over a ``state_dict`` with the reference's keys.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this module, and only
as the checker or the reported baseline.  The product (``floodplanet_code_b200``) never
imports it and fails loudly without its CUDA library.

All arithmetic lives in torch ATen ops, exactly as in the reference (a pure-PyTorch project,
torch pinned at 1.10.2 in environment.yml:114; the op semantics used here are unchanged in the
installed torch 2.11).  Each function cites the reference lines it restates:

  conv3x3+BN+ReLU x2      st_water_seg/models/unet.py:6-20      (DoubleConv)
  maxpool2 + DoubleConv   st_water_seg/models/unet.py:23-32     (Down)
  up x2 + pad + cat + DC  st_water_seg/models/unet.py:35-67     (Up, bilinear branch)
  1x1 head                st_water_seg/models/unet.py:70-77     (OutConv)
  wiring                  st_water_seg/models/unet.py:80-111    (UNet.__init__/forward)
  early-fusion concat     st_water_seg/models/ef_model.py:24-47
  encode / decode halves  st_water_seg/models/unet.py:113-131, 134-191 (UNetEncoder / UNetDecoder)
  late fusion             st_water_seg/models/lf_model.py:29-92
  loss / NaN guard / pred st_water_seg/models/water_seg_model.py:40,98-107
  Adam                    st_water_seg/models/water_seg_model.py:198-205

Pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against the REFERENCE ITSELF, imported by file path in the build container, by
``tests/golden/make_golden.py``; the resulting fixtures are committed under ``tests/golden/``
and re-checked by ``tests/test_oracle.py`` (CPU).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

EXTRA_KEYS = ('dem', 'slope', 'preflood', 'pre_post_difference', 'hand')  # ef_model.py:27-41

# (prefix of the DoubleConv's nn.Sequential, cin, mid, cout)  -- unet.py:88-97 with bilinear=True
def _double_convs(n_channels: int) -> List[Tuple[str, int, int, int]]:
    return [
        ("inc.double_conv", n_channels, 64, 64),
        ("down1.maxpool_conv.1.double_conv", 64, 128, 128),
        ("down2.maxpool_conv.1.double_conv", 128, 256, 256),
        ("down3.maxpool_conv.1.double_conv", 256, 512, 512),
        ("down4.maxpool_conv.1.double_conv", 512, 512, 512),
        ("up1.conv.double_conv", 1024, 512, 256),
        ("up2.conv.double_conv", 512, 256, 128),
        ("up3.conv.double_conv", 256, 128, 64),
        ("up4.conv.double_conv", 128, 64, 64),
    ]


def init_state_dict(n_channels: int, n_classes: int, seed: Optional[int] = 0) -> "OrderedDict[str, torch.Tensor]":
    """Fresh parameters with the reference's default initialisation, consuming the RNG in the
    same order as ``UNet.__init__`` (unet.py:82-98): per DoubleConv conv/bn/conv/bn, then the
    head -- so ``torch.manual_seed(s); UNet(...)`` in the reference gives identical tensors."""
    if seed is not None:
        torch.manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for prefix, cin, mid, cout in _double_convs(n_channels):
        for idx, (ci, co) in ((0, (cin, mid)), (3, (mid, cout))):
            conv = nn.Conv2d(ci, co, kernel_size=3, padding=1)
            bn = nn.BatchNorm2d(co)
            sd[f"{prefix}.{idx}.weight"] = conv.weight.detach().clone()
            sd[f"{prefix}.{idx}.bias"] = conv.bias.detach().clone()
            for k, v in bn.state_dict().items():
                sd[f"{prefix}.{idx + 1}.{k}"] = v.detach().clone()
    head = nn.Conv2d(64, n_classes, kernel_size=1)
    sd["outc.conv.weight"] = head.weight.detach().clone()
    sd["outc.conv.bias"] = head.bias.detach().clone()
    return sd


def trainable_keys(sd: Dict[str, torch.Tensor]) -> List[str]:
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


def _bf16(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 and back (straight-through for autograd: d round(t)/dt := 1)."""
    return t + (t.detach().to(torch.bfloat16).to(torch.float32) - t.detach())


def _conv_bn_relu_bf16(x, sd, conv, bn, training: bool):
    """Same layer with the CUDA path's storage precision emulated (NOT the reference's
    arithmetic -- a diagnostic that separates kernel bugs from bf16 rounding): bf16 conv
    operands, fp32 accumulation, the conv output stored in bf16, BatchNorm statistics and
    normalisation taken from those stored values, bf16-stored activation.  The conv bias only
    shifts the batch mean, so it is kept out of the GEMM exactly like the kernels do.
    Rounding is straight-through, so autograd gives the gradients an exact-arithmetic backward
    pass would produce AT THE SAME (bf16) activations and ReLU masks."""
    y = _bf16(F.conv2d(x, _bf16(sd[f"{conv}.weight"]), None, padding=1))
    if training:
        sd[f"{bn}.num_batches_tracked"] += 1
        with torch.no_grad():
            mean = y.mean((0, 2, 3))
            var = y.var((0, 2, 3), unbiased=False)
            n = y.numel() / y.shape[1]
            sd[f"{bn}.running_mean"].mul_(0.9).add_(0.1 * (mean + sd[f"{conv}.bias"]))
            sd[f"{bn}.running_var"].mul_(0.9).add_(0.1 * var * n / max(n - 1, 1))
        out = F.batch_norm(y, None, None, sd[f"{bn}.weight"], sd[f"{bn}.bias"], True, 0.1, 1e-5)
    else:
        scale = sd[f"{bn}.weight"] / torch.sqrt(sd[f"{bn}.running_var"] + 1e-5)
        shift = sd[f"{bn}.bias"] + (sd[f"{conv}.bias"] - sd[f"{bn}.running_mean"]) * scale
        out = y * scale[None, :, None, None] + shift[None, :, None, None]
    return _bf16(F.relu(out))


_EMULATE_BF16 = False


def _conv_bn_relu(x, sd, conv, bn, training: bool):
    """unet.py:14-17 -- Conv2d(k=3,p=1,bias) -> BatchNorm2d(eps=1e-5, momentum=0.1) -> ReLU."""
    if _EMULATE_BF16:
        return _conv_bn_relu_bf16(x, sd, conv, bn, training)
    x = F.conv2d(x, sd[f"{conv}.weight"], sd[f"{conv}.bias"], padding=1)
    if training:
        sd[f"{bn}.num_batches_tracked"] += 1
    x = F.batch_norm(x, sd[f"{bn}.running_mean"], sd[f"{bn}.running_var"], sd[f"{bn}.weight"],
                     sd[f"{bn}.bias"], training, 0.1, 1e-5)
    return F.relu(x)


def _double_conv(x, sd, prefix, training):
    x = _conv_bn_relu(x, sd, f"{prefix}.0", f"{prefix}.1", training)
    return _conv_bn_relu(x, sd, f"{prefix}.3", f"{prefix}.4", training)


def _up(x1, x2, sd, prefix, training):
    """unet.py:54-67 -- bilinear x2 (align_corners=True), zero-pad to the skip size with the
    smaller half on the left/top, cat([skip, upsampled]) along C, DoubleConv."""
    x1 = F.interpolate(x1, scale_factor=2, mode='bilinear', align_corners=True)
    dy = x2.size(2) - x1.size(2)
    dx = x2.size(3) - x1.size(3)
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    if _EMULATE_BF16:
        x1 = _bf16(x1)
    return _double_conv(torch.cat([x2, x1], dim=1), sd, prefix, training)


def unet_forward_bf16_emulated(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = True):
    """Diagnostic variant of :func:`unet_forward` with the CUDA path's bf16 storage points
    emulated (see :func:`_conv_bn_relu_bf16`).  Diagnostic only: two bf16 implementations of this
    network differ from EACH OTHER by 2-4 % at random init (measured: CUDA path vs this 1.6-3.1 %, stock
    torch.autocast vs this 4-5 %), because where exactly the roundings fall differs and the network
    amplifies them; tests assert < 4 % here.  The per-step parity at 1e-2 / 2e-2 is the teacher-forced
    walk (tests/test_teacher_forced_gpu.py), the end-to-end bound is the measured autocast envelope
    (tests/golden/autocast_envelope.json)."""
    global _EMULATE_BF16
    _EMULATE_BF16 = True
    try:
        return unet_forward(sd, _bf16(x), training)
    finally:
        _EMULATE_BF16 = False


def unet_encode(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = True, prefix: str = ""):
    """unet.py:113-120 (UNet.encode) / :150-159 (UNetEncoder.forward) -> [x1..x5]."""
    x1 = _double_conv(x, sd, f"{prefix}inc.double_conv", training)
    x2 = _double_conv(F.max_pool2d(x1, 2), sd, f"{prefix}down1.maxpool_conv.1.double_conv", training)
    x3 = _double_conv(F.max_pool2d(x2, 2), sd, f"{prefix}down2.maxpool_conv.1.double_conv", training)
    x4 = _double_conv(F.max_pool2d(x3, 2), sd, f"{prefix}down3.maxpool_conv.1.double_conv", training)
    x5 = _double_conv(F.max_pool2d(x4, 2), sd, f"{prefix}down4.maxpool_conv.1.double_conv", training)
    return [x1, x2, x3, x4, x5]


def unet_decode(sd: Dict[str, torch.Tensor], feats: Sequence[torch.Tensor], training: bool = True,
                prefix: str = "", head: bool = True):
    """unet.py:122-131 (UNet.decode) / :176-191 (UNetDecoder.forward, get_output_feats)."""
    x1, x2, x3, x4, x5 = feats
    u = _up(x5, x4, sd, f"{prefix}up1.conv.double_conv", training)
    u = _up(u, x3, sd, f"{prefix}up2.conv.double_conv", training)
    u = _up(u, x2, sd, f"{prefix}up3.conv.double_conv", training)
    u = _up(u, x1, sd, f"{prefix}up4.conv.double_conv", training)
    if not head:
        return u
    return F.conv2d(u, sd[f"{prefix}outc.conv.weight"], sd[f"{prefix}outc.conv.bias"])


def unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = True,
                 return_features: bool = False):
    """unet.py:100-111.  ``sd`` BN buffers are updated in place when ``training``."""
    feats = unet_encode(sd, x, training)
    logits = unet_decode(sd, feats, training)
    if return_features:
        return logits, feats
    return logits


# ---------------------------------------------------------------------------------------------
# late fusion (lf_model.py)
# ---------------------------------------------------------------------------------------------
LF_BATCH_KEYS = (('image', 'ms_image'), ('dem', 'dem'), ('slope', 'slope'), ('preflood', 'preflood'),
                 ('pre_post_difference', 'pre_post_difference'), ('hand', 'hand'))   # lf_model.py:60-81
LF_FEAT_SIZES = (64, 128, 256, 512, 512)                                             # lf_model.py:42


def _init_double_conv(sd, prefix, cin, mid, cout):
    for idx, (ci, co) in ((0, (cin, mid)), (3, (mid, cout))):
        conv = nn.Conv2d(ci, co, kernel_size=3, padding=1)
        bn = nn.BatchNorm2d(co)
        sd[f"{prefix}.{idx}.weight"] = conv.weight.detach().clone()
        sd[f"{prefix}.{idx}.bias"] = conv.bias.detach().clone()
        for k, v in bn.state_dict().items():
            sd[f"{prefix}.{idx + 1}.{k}"] = v.detach().clone()


def init_lf_state_dict(in_channels: Dict[str, int], n_classes: int,
                       seed: Optional[int] = 0) -> "OrderedDict[str, torch.Tensor]":
    """Fresh late-fusion parameters, consuming the RNG in the order of
    ``LateFusionModel._build_model`` (lf_model.py:29-45): one UNetEncoder per input in dict
    order, the UNetDecoder (up1..up4, outc), then the five concat convs."""
    if seed is not None:
        torch.manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    dcs = _double_convs(0)
    for name, c in in_channels.items():
        for prefix, cin, mid, cout in dcs[:5]:
            _init_double_conv(sd, f"encoders.{name}.{prefix}", c if prefix.startswith("inc") else cin, mid, cout)
    for prefix, cin, mid, cout in dcs[5:]:
        _init_double_conv(sd, f"decoder.{prefix}", cin, mid, cout)
    head = nn.Conv2d(64, n_classes, kernel_size=1)
    sd["decoder.outc.conv.weight"] = head.weight.detach().clone()
    sd["decoder.outc.conv.bias"] = head.bias.detach().clone()
    for i, fs in enumerate(LF_FEAT_SIZES):
        cc = nn.Conv2d(fs * len(in_channels), fs, 1, 1)
        sd[f"concat_convs.{i}.weight"] = cc.weight.detach().clone()
        sd[f"concat_convs.{i}.bias"] = cc.bias.detach().clone()
    return sd


def lf_forward(sd: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], training: bool = True):
    """lf_model.py:58-92 -- per-modality encoders, per-level torch.concat in the fixed key
    order, 1x1 concat_convs, decoder."""
    feats = None
    for bkey, enc in LF_BATCH_KEYS:
        if bkey != 'image' and bkey not in batch:
            continue
        f = unet_encode(sd, batch[bkey], training, prefix=f"encoders.{enc}.")
        feats = f if feats is None else [torch.concat([a, b], dim=1) for a, b in zip(feats, f)]
    fused = []
    for i, f in enumerate(feats):
        if _EMULATE_BF16:
            fused.append(_bf16(F.conv2d(f, _bf16(sd[f"concat_convs.{i}.weight"]), sd[f"concat_convs.{i}.bias"])))
        else:
            fused.append(F.conv2d(f, sd[f"concat_convs.{i}.weight"], sd[f"concat_convs.{i}.bias"]))
    return unet_decode(sd, fused, training, prefix="decoder.")


def lf_forward_bf16_emulated(sd, batch, training: bool = True):
    """Late-fusion forward with the CUDA path's bf16 storage points emulated (diagnostic)."""
    global _EMULATE_BF16
    _EMULATE_BF16 = True
    try:
        b = {k: (_bf16(v) if v.is_floating_point() else v) for k, v in batch.items()}
        return lf_forward(sd, b, training)
    finally:
        _EMULATE_BF16 = False


def lf_training_step(sd: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor],
                     ignore_index: Optional[int], emulate_bf16: bool = False):
    """One late-fusion training step (water_seg_model.py:98-136 with lf_model.forward)."""
    keys = trainable_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
        sd[k].grad = None
    logits = lf_forward_bf16_emulated(sd, batch, True) if emulate_bf16 else lf_forward(sd, batch, True)
    loss, pred = masked_ce(logits, batch['target'], ignore_index)
    loss.backward()
    grads = {k: (sd[k].grad.detach().clone() if sd[k].grad is not None else torch.zeros_like(sd[k]))
             for k in keys}
    for k in keys:
        sd[k].requires_grad_(False)
        sd[k].grad = None
    return loss.detach(), pred, logits.detach(), grads


def early_fusion_input(batch: Dict[str, torch.Tensor]) -> torch.Tensor:
    """ef_model.py:24-44 -- image then dem, slope, preflood, pre_post_difference, hand."""
    images = batch['image']
    for k in EXTRA_KEYS:
        if k in batch:
            images = torch.concat([images, batch[k]], dim=1)
    return images


def masked_ce(logits: torch.Tensor, target: torch.Tensor, ignore_index: Optional[int]):
    """water_seg_model.py:40,103-107 -- CE(mean over non-ignored), NaN -> 0, argmax(dim=1)."""
    ii = -100 if ignore_index is None else ignore_index
    loss = F.cross_entropy(logits, target, ignore_index=ii)
    if torch.isnan(loss):
        loss = torch.nan_to_num(loss)
    return loss, logits.argmax(dim=1)


def training_step(sd: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor],
                  ignore_index: Optional[int], early_fusion: bool = True, emulate_bf16: bool = False):
    """Forward + loss + backward of one reference training step (water_seg_model.py:98-136 and
    the ``loss.backward()`` Lightning runs).  Returns (loss, pred, logits, {name: grad}).
    ``emulate_bf16`` switches to the storage-precision diagnostic (see _conv_bn_relu_bf16)."""
    keys = trainable_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
        sd[k].grad = None
    x = early_fusion_input(batch) if early_fusion else batch['image']
    logits = unet_forward_bf16_emulated(sd, x, True) if emulate_bf16 else unet_forward(sd, x, training=True)
    loss, pred = masked_ce(logits, batch['target'], ignore_index)
    loss.backward()
    grads = {k: (sd[k].grad.detach().clone() if sd[k].grad is not None else torch.zeros_like(sd[k]))
             for k in keys}
    for k in keys:
        sd[k].requires_grad_(False)
        sd[k].grad = None
    return loss.detach(), pred, logits.detach(), grads


def confusion_counts(pred: torch.Tensor, target: torch.Tensor, n_classes: int,
                     ignore_index: Optional[int]) -> torch.Tensor:
    keep = torch.ones_like(target, dtype=torch.bool) if ignore_index is None else target != ignore_index
    idx = target[keep] * n_classes + pred[keep]
    return torch.bincount(idx, minlength=n_classes * n_classes).view(n_classes, n_classes)


def micro_metrics(conf: torch.Tensor, ignore_index: Optional[int] = None) -> Dict[str, float]:
    """torchmetrics multiclass micro F1 / Jaccard / Accuracy with ignore_index
    (water_seg_model.py:46-63) expressed on confusion counts (rows = target, ignored targets
    already removed).  torchmetrics (>= 0.11, implied by the reference's ``task="multiclass"``
    arguments; not vendored in /root/reference, not installed here) is restated from its published
    algorithm: micro F1 = Accuracy = tp / (tp + fn); Jaccard per class denom_c = colsum_c + rowsum_c -
    diag_c, micro = sum(diag) / (sum(denom) - denom[ignore_index]) when 0 <= ignore_index < C
    (functional/classification/jaccard.py, `_jaccard_index_reduce`).  Written out per class here on
    purpose, unlike the closed form the product uses.  PARITY UNPINNED for this one function: no
    torchmetrics build is available to generate a fixture."""
    conf = conf.double()
    c = conf.shape[0]
    diag = conf.diagonal()
    total = conf.sum()
    acc = float(torch.nan_to_num(diag.sum() / total))
    denom = conf.sum(0) + conf.sum(1) - diag
    dsum = denom.sum()
    if ignore_index is not None and 0 <= ignore_index < c:
        dsum = dsum - denom[ignore_index]
    return {"F1": acc, "Accuracy": acc, "Jaccard": float(torch.nan_to_num(diag.sum() / dsum))}


def synthetic_batch(n: int, c: int, h: int, w: int, seed: int = 0, ignore_frac: float = 0.58,
                    block: int = 32, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """image ~ U[0,1) f32; int64 target, spatially blocky, ~58% class 0
    (ignored under the default config) / 42% class 1, never class 2."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    image = torch.rand(n, c, h, w, generator=g)
    hb, wb = (h + block - 1) // block, (w + block - 1) // block
    coarse = torch.rand(n, 1, hb, wb, generator=g)
    field = F.interpolate(coarse, size=(hb * block, wb * block), mode="nearest")[:, 0, :h, :w]
    target = (field >= ignore_frac).long()
    return {"image": image.to(device), "target": target.to(device)}
