"""Drop-in ``LateFusionModel`` (reference: st_water_seg/models/lf_model.py:9-92).

Same constructor, module tree (``encoders`` ModuleDict of UNetEncoder keyed by input name,
``decoder`` UNetDecoder, ``concat_convs`` ModuleList of five ``nn.Conv2d(fs*k, fs, 1, 1)``) and
therefore the same ``state_dict`` as the reference; ``forward`` hands every tensor to
:class:`~floodplanet_code_b200.engine.LateFusionEngine`: per-modality encoders write straight
into channel slices of one fusion buffer per level (no ``torch.concat``), the 1x1 fusion convs
run on the tensor cores, then the ordinary decoder.  One autograd node for the whole network.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .engine import LateFusionEngine
from .unet import UNetDecoder, UNetEncoder, _check_cuda_nchw
from .water_seg_model import MODELS, WaterSegmentationModel


class _LateFusionFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, module, keys, *tensors):
        n_images = len(keys)
        engine = module._engine
        images = dict(zip(keys, tensors[:n_images]))
        plist = tensors[n_images:]
        params = dict(zip(engine.names, plist))
        logits, st = engine.forward(images, params, module._engine_buffers(), training=module.decoder.training,
                                    save=True)
        ctx.engine, ctx.state, ctx.n_images = engine, st, n_images
        ctx.save_for_backward(*plist)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        engine = ctx.engine
        if ctx.state is None:
            raise RuntimeError("floodplanet_b200: backward called twice on the same late-fusion forward")
        params = dict(zip(engine.names, ctx.saved_tensors))
        grads, _slab = engine.backward(ctx.state, dlogits, params)
        ctx.state = None
        return (None, None) + (None,) * ctx.n_images + tuple(grads[n] for n in engine.names)


class LateFusionModel(WaterSegmentationModel):

    # batch key -> encoder name, in the concatenation order of lf_model.py:60-81
    BATCH_KEYS = (('image', 'ms_image'), ('dem', 'dem'), ('slope', 'slope'), ('preflood', 'preflood'),
                  ('pre_post_difference', 'pre_post_difference'), ('hand', 'hand'))

    def __init__(self,
                 in_channels,
                 n_classes,
                 lr,
                 log_image_iter=50,
                 to_rgb_fcn=None,
                 ignore_index=None,
                 optimizer_name='adam',
                 feat_fusion='concat_conv'):
        self.feat_fusion = feat_fusion
        super().__init__(in_channels,
                         n_classes,
                         lr,
                         log_image_iter=log_image_iter,
                         to_rgb_fcn=to_rgb_fcn,
                         optimizer_name=optimizer_name,
                         ignore_index=ignore_index)

    def _build_model(self):
        # reference :29-45
        self.encoders = nn.ModuleDict()
        if type(self.in_channels) is dict:
            for input_name, feature_channels in self.in_channels.items():
                self.encoders[input_name] = UNetEncoder(feature_channels)

        self.decoder = UNetDecoder(self.n_classes)

        if self.feat_fusion == 'concat_conv':
            unet_feat_sizes = [64, 128, 256, 512, 512]
            self.concat_convs = nn.ModuleList()
            for fs in unet_feat_sizes:
                self.concat_convs.append(nn.Conv2d(fs * len(self.in_channels), fs, 1, 1))
        self._engine = LateFusionEngine(dict(self.in_channels), self.n_classes) \
            if type(self.in_channels) is dict and self.feat_fusion == 'concat_conv' else None

    def _set_model_to_train(self):
        self.encoders.train()
        self.decoder.train()
        self.concat_convs.train()

    def _set_model_to_eval(self):
        self.encoders.eval()
        self.decoder.eval()
        self.concat_convs.eval()

    def _engine_params(self):
        own = dict(self.named_parameters())
        return {n: own[n] for n in self._engine.names}

    def _engine_buffers(self):
        return dict(self.named_buffers())

    def forward(self, batch):
        if self.feat_fusion != 'concat_conv':
            raise NotImplementedError
        keys, images = [], []
        batch_keys = list(batch.keys())
        for bkey, enc in self.BATCH_KEYS:
            if bkey == 'image' or bkey in batch_keys:
                self.encoders[enc]          # KeyError for a modality without an encoder, as in the reference
                keys.append(enc)
                images.append(batch[bkey])
        _check_cuda_nchw(images)
        engine = self._engine
        params = self._engine_params()
        training = self.decoder.training
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params.values())
        if needs_grad:
            return _LateFusionFunction.apply(self, keys, *images, *[params[n] for n in engine.names])
        with torch.no_grad():
            logits, _ = engine.forward(dict(zip(keys, images)), params, self._engine_buffers(),
                                       training=training, save=False)
        return logits

    @property
    def kernel_launches(self) -> int:
        return self._engine.launches


MODELS['lf_model'] = LateFusionModel
