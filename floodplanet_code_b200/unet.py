"""Drop-in ``UNet`` (reference: st_water_seg/models/unet.py).

Same constructor, attributes, module tree and therefore the same ``state_dict`` keys, shapes
and dtypes as the reference (`inc.double_conv.{0,1,3,4}.*`, `down{1-4}.maxpool_conv.1.…`,
`up{1-4}.conv.…`, `outc.conv.*`, BN buffers incl. int64 `num_batches_tracked`), so existing
checkpoints load with ``strict=True`` and stock ``torch.optim.Adam`` steps the fp32 master
parameters.  The submodules are *parameter containers only*: ``UNet.forward`` does not call
them -- it hands their tensors to :class:`~floodplanet_code_b200.engine.UNetEngine`, which
runs the sm_100a kernels through the C ABI.  There is no CPU / eager fallback: a non-CUDA
input raises.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from .engine import UNetEngine


class _ContainerOnly(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover - guard rail
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container of the B200 UNet; call UNet.forward "
            "(per-block eager execution is not part of this path)")


class DoubleConv(_ContainerOnly):
    """(convolution => [BN] => ReLU) * 2 -- reference unet.py:6-20."""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        if not mid_channels:
            mid_channels = out_channels
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(mid_channels), nn.ReLU(inplace=True),
            nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))


class Down(_ContainerOnly):
    """Downscaling with maxpool then double conv -- reference unet.py:23-32."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))


class Up(_ContainerOnly):
    """Upscaling then double conv -- reference unet.py:35-67 (bilinear branch only)."""

    def __init__(self, in_channels, out_channels, bilinear=True):
        super().__init__()
        if not bilinear:
            raise NotImplementedError(
                "bilinear=False (ConvTranspose2d) is never selected by the reference "
                "(water_seg_model.py:85) and is not built")
        self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        self.conv = DoubleConv(in_channels, out_channels, in_channels // 2)


class OutConv(_ContainerOnly):
    """1x1 classifier head -- reference unet.py:70-77."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)


class _UNetFunction(torch.autograd.Function):
    """One autograd node for the whole network: forward and backward are flat kernel schedules."""

    @staticmethod
    def forward(ctx, module: "UNet", n_images: int, *tensors):
        images = tensors[:n_images]
        plist = tensors[n_images:]
        engine = module._engine
        params = dict(zip(engine.names, plist))
        buffers = dict(module.named_buffers())
        logits, st = engine.forward(images, params, buffers, training=True, save=True)
        ctx.engine = engine
        ctx.state = st
        ctx.n_images = n_images
        ctx.save_for_backward(*plist)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        engine: UNetEngine = ctx.engine
        if ctx.state is None:
            raise RuntimeError("floodplanet_b200: backward called twice on the same UNet forward")
        params = dict(zip(engine.names, ctx.saved_tensors))
        grads, _slab = engine.backward(ctx.state, dlogits, params)
        ctx.state = None  # release saved activations
        return (None, None) + (None,) * ctx.n_images + tuple(grads[n] for n in engine.names)


class UNet(nn.Module):
    """``UNet(n_channels, n_classes, bilinear=True)`` -- reference unet.py:80-131."""

    def __init__(self, n_channels, n_classes, bilinear=True):
        super(UNet, self).__init__()
        self.n_channels = n_channels
        self.n_classes = n_classes
        self.bilinear = bilinear

        self.inc = DoubleConv(n_channels, 64)
        self.down1 = Down(64, 128)
        self.down2 = Down(128, 256)
        self.down3 = Down(256, 512)
        factor = 2 if bilinear else 1
        self.down4 = Down(512, 1024 // factor)
        self.up1 = Up(1024, 512 // factor, bilinear)
        self.up2 = Up(512, 256 // factor, bilinear)
        self.up3 = Up(256, 128 // factor, bilinear)
        self.up4 = Up(128, 64, bilinear)
        self.outc = OutConv(64, n_classes)

        self._engine = UNetEngine(n_channels, n_classes)

    # -- the hot path -----------------------------------------------------------------------
    def forward(self, x):
        return self.forward_fused([x])

    def forward_fused(self, images: Sequence[torch.Tensor]):
        """Forward on several NCHW tensors that are to be concatenated along C (early fusion,
        ef_model.py:24-47) without materialising the concatenation."""
        images = list(images)
        for t in images:
            if not t.is_cuda:
                raise RuntimeError(
                    "floodplanet_b200.UNet runs on CUDA (sm_100a) only; got a "
                    f"{t.device.type} tensor and there is no CPU fallback")
            if t.dim() != 4:
                raise RuntimeError(f"expected NCHW input, got shape {tuple(t.shape)}")
        engine = self._engine
        params = dict(self.named_parameters())
        needs_grad = (torch.is_grad_enabled() and self.training
                      and any(p.requires_grad for p in params.values()))
        if needs_grad:
            return _UNetFunction.apply(self, len(images), *images, *[params[n] for n in engine.names])
        with torch.no_grad():
            logits, _ = engine.forward(images, params, dict(self.named_buffers()),
                                       training=self.training, save=False)
        return logits

    # -- reference API kept for completeness (late fusion only; out of the hot path) ----------
    def encode(self, x):
        raise NotImplementedError(
            "UNet.encode/decode are only used by the late-fusion model (lf_model.py), which is "
            "outside the B200 hot-path scope (SURVEY.md section 8f)")

    def decode(self, feats):
        raise NotImplementedError(
            "UNet.encode/decode are only used by the late-fusion model (lf_model.py), which is "
            "outside the B200 hot-path scope (SURVEY.md section 8f)")

    @property
    def kernel_launches(self) -> int:
        """CUDA kernels launched by the most recent forward or backward pass."""
        return self._engine.launches
