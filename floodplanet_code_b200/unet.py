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

from . import ops
from .engine import DecoderEngine, EncoderEngine, UNetEngine


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
        # module.training False = eval-mode (running-statistics) BatchNorm that stays differentiable
        logits, st = engine.forward(images, params, buffers, training=module.training, save=True)
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
        self._enc_engine = None   # built on first encode() / decode()
        self._dec_engine = None

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
            if t.requires_grad and torch.is_grad_enabled():
                raise RuntimeError(
                    "floodplanet_b200.UNet does not produce gradients with respect to its input images "
                    "(the first convolution has no dgrad on this path; the reference's training / "
                    "inference scripts never ask for one) -- detach the input")
        engine = self._engine
        params = dict(self.named_parameters())
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params.values())
        if needs_grad:
            # train(): batch-statistics BatchNorm; eval() with grad mode on: the same differentiable schedule
            # with the running statistics as constants (frozen-BatchNorm fine-tuning; the reference's modules
            # support autograd in eval mode too).  Validation / infer / predict run under no_grad -> below.
            return _UNetFunction.apply(self, len(images), *images, *[params[n] for n in engine.names])
        with torch.no_grad():
            logits, _ = engine.forward(images, params, dict(self.named_buffers()),
                                       training=self.training, save=False)
        return logits

    # -- feature-level API (reference unet.py:113-131; used by late fusion) ---------------------
    def encode(self, x):
        """[x1, x2, x3, x4, x5] as fp32 NCHW tensors (unet.py:113-120)."""
        if self._enc_engine is None:
            self._enc_engine = EncoderEngine(self.n_channels)
        return _run_encoder_module(self, self._enc_engine, [x])

    def decode(self, feats):
        """logits from the five features (unet.py:122-131)."""
        if self._dec_engine is None:
            self._dec_engine = DecoderEngine(self.n_classes)
        return _run_decoder_module(self, self._dec_engine, feats)

    @property
    def kernel_launches(self) -> int:
        """CUDA kernels launched by the most recent forward or backward pass."""
        return self._engine.launches


# ---------------------------------------------------------------------------------------------
# encoder / decoder halves as stand-alone autograd nodes (fp32 NCHW feature lists at the seam)
# ---------------------------------------------------------------------------------------------
def _check_cuda_nchw(tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(
                "floodplanet_b200 runs on CUDA (sm_100a) only; got a "
                f"{t.device.type} tensor and there is no CPU fallback")
        if t.dim() != 4:
            raise RuntimeError(f"expected NCHW input, got shape {tuple(t.shape)}")


class _EncoderFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, module, engine, n_images, *tensors):
        images, plist = tensors[:n_images], tensors[n_images:]
        params = dict(zip(engine.names, plist))
        feats, st = engine.forward(images, params, dict(module.named_buffers()), training=module.training, save=True)
        ctx.engine, ctx.state, ctx.n_images = engine, st, n_images
        ctx.save_for_backward(*plist)
        return tuple(ops.nhwc_bf16_to_nchw_f32(f) for f in feats)

    @staticmethod
    def backward(ctx, *d_feats):
        engine = ctx.engine
        if ctx.state is None:
            raise RuntimeError("floodplanet_b200: backward called twice on the same encoder forward")
        st = ctx.state
        params = dict(zip(engine.names, ctx.saved_tensors))
        d_nhwc = []
        for l, g in enumerate(d_feats):
            hh, ww = st.sizes[l]
            c = (64, 128, 256, 512, 512)[l]
            buf = torch.zeros((st.n, hh, ww, c), dtype=torch.bfloat16, device=ctx.saved_tensors[0].device)
            if g is not None:
                ops.nchw_f32_to_nhwc_bf16(g, buf)
            d_nhwc.append(buf)
        grads, _ = engine.backward(st, d_nhwc, params)
        ctx.state = None
        return (None, None, None) + (None,) * ctx.n_images + tuple(grads[n] for n in engine.names)


class _DecoderFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, module, engine, *tensors):
        feats, plist = tensors[:5], tensors[5:]
        params = dict(zip(engine.names, plist))
        logits, st = engine.forward(feats, params, dict(module.named_buffers()), training=module.training, save=True)
        ctx.engine, ctx.state = engine, st
        ctx.save_for_backward(*plist)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        engine = ctx.engine
        if ctx.state is None:
            raise RuntimeError("floodplanet_b200: backward called twice on the same decoder forward")
        params = dict(zip(engine.names, ctx.saved_tensors))
        grads, _, d_feats = engine.backward(ctx.state, dlogits, params)
        ctx.state = None
        d_nchw = tuple(ops.nhwc_bf16_to_nchw_f32(g) if ctx.needs_input_grad[2 + l] else None
                       for l, g in enumerate(d_feats))
        return (None, None) + d_nchw + tuple(grads[n] for n in engine.names)


def _run_encoder_module(module: nn.Module, engine: EncoderEngine, images):
    images = list(images)
    _check_cuda_nchw(images)
    params = dict(module.named_parameters())
    plist = [params[n] for n in engine.names]
    if torch.is_grad_enabled() and any(p.requires_grad for p in plist):
        return list(_EncoderFunction.apply(module, engine, len(images), *images, *plist))
    with torch.no_grad():
        feats, _ = engine.forward(images, params, dict(module.named_buffers()),
                                  training=module.training, save=False)
        return [ops.nhwc_bf16_to_nchw_f32(f) for f in feats]


def _run_decoder_module(module: nn.Module, engine: DecoderEngine, feats, head: bool = True):
    feats = list(feats)
    _check_cuda_nchw(feats)
    params = dict(module.named_parameters())
    plist = [params[n] for n in engine.names]
    needs_grad = torch.is_grad_enabled() and (
        any(p.requires_grad for p in plist) or any(f.requires_grad for f in feats))
    if needs_grad:
        if not head:
            raise RuntimeError("floodplanet_b200: get_output_feats is inference-only on this path "
                               "(no caller in the reference differentiates through it)")
        return _DecoderFunction.apply(module, engine, *feats, *plist)
    with torch.no_grad():
        out, _ = engine.forward(feats, params, dict(module.named_buffers()), training=module.training,
                                save=False, head=head)
        return out if head else ops.nhwc_bf16_to_nchw_f32(out)


class UNetEncoder(nn.Module):
    """``UNetEncoder(n_channels, bilinear=True, base_feat_channels=64)`` -- reference
    unet.py:134-159.  forward(x) -> [x1..x5] (fp32 NCHW)."""

    def __init__(self, n_channels, bilinear=True, base_feat_channels=64):
        super(UNetEncoder, self).__init__()
        if base_feat_channels != 64 or not bilinear:
            raise NotImplementedError("the B200 path builds the widths the reference instantiates "
                                      "(base_feat_channels=64, bilinear=True; lf_model.py:36-38)")
        self.n_channels = n_channels
        self.bilinear = bilinear
        bfc = base_feat_channels
        self.base_feat_channels = base_feat_channels

        self.inc = DoubleConv(n_channels, bfc)
        self.down1 = Down(bfc, bfc * 2)
        self.down2 = Down(bfc * 2, bfc * 4)
        self.down3 = Down(bfc * 4, bfc * 8)
        factor = 2 if bilinear else 1
        self.down4 = Down(bfc * 8, (bfc * 16) // factor)
        self._engine = EncoderEngine(n_channels)

    def forward(self, x):
        return _run_encoder_module(self, self._engine, [x])


class UNetDecoder(nn.Module):
    """``UNetDecoder(n_classes, bilinear=True, channel_factor=1, base_feat_channels=64)`` --
    reference unet.py:162-191."""

    def __init__(self, n_classes, bilinear=True, channel_factor=1, base_feat_channels=64):
        super(UNetDecoder, self).__init__()
        if base_feat_channels != 64 or not bilinear or channel_factor != 1:
            raise NotImplementedError("the B200 path builds the widths the reference instantiates "
                                      "(channel_factor=1, base_feat_channels=64, bilinear=True; "
                                      "lf_model.py:40)")
        self.n_classes = n_classes
        self.bilinear = bilinear
        cf = channel_factor
        bfc = base_feat_channels
        self.base_feat_channels = base_feat_channels

        factor = 2 if bilinear else 1
        self.up1 = Up((bfc * 16) * cf, (bfc * 8) // factor, bilinear)
        self.up2 = Up((bfc * 8) // factor * (cf + 1), (bfc * 4) // factor, bilinear)
        self.up3 = Up((bfc * 4) // factor * (cf + 1), (bfc * 2) // factor, bilinear)
        self.up4 = Up((bfc * 2) // factor * (cf + 1), bfc, bilinear)
        self.outc = OutConv(bfc, n_classes)
        self._engine = DecoderEngine(n_classes)

    def forward(self, feats):
        return _run_decoder_module(self, self._engine, feats)

    def get_output_feats(self, feats):
        return _run_decoder_module(self, self._engine, feats, head=False)
