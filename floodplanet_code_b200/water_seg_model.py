"""Drop-in ``WaterSegmentationModel`` / ``EarlyFusionModel`` LightningModules and the
``MODELS`` / ``build_model`` plugin seam.

Reference: st_water_seg/models/water_seg_model.py (ctor :16-44, forward :87-90,
training_step :98-136, validation_step :138-179, test_step :181-196,
configure_optimizers :198-205), models/ef_model.py:24-47 and models/__init__.py:5-20.
Same constructor signatures, attribute names and step semantics; what the steps dispatch
to is the B200 engine (UNet kernels + fused masked-CE/argmax/confusion kernel).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn
import torch.optim as optim

from .loss import MaskedCrossEntropyLoss
from .metrics import MicroSegmentationMetrics
from .unet import UNet

try:  # pytorch_lightning is optional in this image; the reference's fit.py needs it
    import pytorch_lightning as pl  # type: ignore
    _LightningBase = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # pragma: no cover - exercised when lightning is absent
    HAVE_LIGHTNING = False

    class _LightningBase(nn.Module):
        """Minimal stand-in exposing what the reference's modules and scripts touch."""

        def __init__(self):
            super().__init__()
            self.logged: Dict[str, torch.Tensor] = {}
            self.current_epoch = 0
            self.global_step = 0
            self.logger = None

        def log_dict(self, metrics, *args, **kwargs):
            self.logged.update(metrics)

        def log(self, name, value, *args, **kwargs):
            self.logged[name] = value

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            model = cls(**kwargs)
            model.load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt)
            return model


class WaterSegmentationModel(_LightningBase):

    def __init__(self,
                 in_channels,
                 n_classes,
                 lr,
                 log_image_iter=50,
                 to_rgb_fcn=None,
                 ignore_index=None,
                 optimizer_name='adam'):
        super().__init__()
        self.lr = lr
        self.n_classes = n_classes
        self.in_channels = in_channels
        self.ignore_index = ignore_index
        self.optimizer_name = optimizer_name

        # Build model.
        self._build_model()

        # Get metrics.
        if self.ignore_index == -1:
            self.ignore_index = self.n_classes - 1
        self.tracked_metrics = self._get_tracked_metrics()

        # Get loss function (fused masked CE + argmax + confusion counts).
        self.loss_func = MaskedCrossEntropyLoss(ignore_index=self.ignore_index)

        # Log images hyperparamters.
        self.to_rgb_fcn = to_rgb_fcn
        self.log_image_iter = log_image_iter

    def _get_tracked_metrics(self, average_mode='micro'):
        metrics = MicroSegmentationMetrics(self.n_classes, self.ignore_index)
        self.train_metrics = metrics.clone(prefix='train_')
        self.valid_metrics = metrics.clone(prefix='val_')
        self.test_metrics = metrics.clone(prefix='test_')

    def _build_model(self):
        # reference :79-85 -- in_channels must be a dict {feature: n_channels}
        if type(self.in_channels) is dict:
            n_in_channels = 0
            for feature_channels in self.in_channels.values():
                n_in_channels += feature_channels
        self.model = UNet(n_in_channels, self.n_classes)

    def forward(self, batch):
        images = batch['image']
        output = self.model(images)
        return output

    def _set_model_to_train(self):
        self.model.train()

    def _set_model_to_eval(self):
        self.model.eval()

    # -- steps ---------------------------------------------------------------------------------
    def _loss_pred_confusion(self, output, target):
        loss = self.loss_func(output, target)
        # reference :104-106: an all-ignored batch gives NaN -> 0 (with zero gradients).  Done
        # on the device without the host sync `if torch.isnan(loss)` would cost.
        loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
        return loss, self.loss_func.last_pred, self.loss_func.last_confusion

    def training_step(self, batch, batch_idx):
        self._set_model_to_train()
        target = batch['target']
        output = self.forward(batch)
        loss, pred, conf = self._loss_pred_confusion(output, target)
        metric_output = self.train_metrics.forward_from_confusion(conf)
        self.log_dict(metric_output, prog_bar=True, on_step=True, on_epoch=True)
        return loss

    def validation_step(self, batch, batch_idx):
        self._set_model_to_eval()
        target = batch['target']
        output = self.forward(batch)
        loss, pred, conf = self._loss_pred_confusion(output, target)
        metric_output = self.valid_metrics.forward_from_confusion(conf)
        self.valid_metrics.update_from_confusion(conf)  # reference quirk :150-151 (double count)
        metric_output['valid_loss'] = loss
        self.log_dict(metric_output, prog_bar=True, on_step=True, on_epoch=True)

    def test_step(self, batch, batch_idx):
        self._set_model_to_eval()
        output = self.forward(batch)
        loss = self.loss_func(output, batch['target'])  # reference :185: no NaN guard here
        self.test_metrics.update_from_confusion(self.loss_func.last_confusion)
        self.log_dict({'test_loss': loss}, prog_bar=True, on_step=True, on_epoch=True)

    def configure_optimizers(self):
        if self.optimizer_name == 'adam':
            optimizer = optim.Adam(self.parameters(), lr=self.lr)
        else:
            raise NotImplementedError(
                f'No implementation for optimizer of name: {self.optimizer_name}')
        return optimizer

    def validation_epoch_end(self, validation_step_outputs):
        if len(validation_step_outputs) == 0:
            self.test_f1_score = 0
            self.test_iou = 0
            self.test_acc = 0
        else:
            metric_output = self.valid_metrics.compute()
            self.log_dict(metric_output)

    def test_epoch_end(self, test_step_outputs) -> None:
        if len(test_step_outputs) == 0:
            pass
        else:
            metric_output = self.test_metrics.compute()
            self.log_dict(metric_output)
            self.f1_score = metric_output['test_MulticlassF1Score'].item()
            self.acc = metric_output['test_MulticlassAccuracy'].item()
            self.iou = metric_output['test_MulticlassJaccardIndex'].item()


class EarlyFusionModel(WaterSegmentationModel):
    """Early fusion: extra rasters are concatenated to the image along C in the fixed order
    dem, slope, preflood, pre_post_difference, hand (reference ef_model.py:24-47).  The concat
    is folded into the NCHW->NHWC ingest kernel instead of a chain of torch.concat copies."""

    EXTRA_KEYS = ('dem', 'slope', 'preflood', 'pre_post_difference', 'hand')

    def __init__(self,
                 in_channels,
                 n_classes,
                 lr,
                 log_image_iter=50,
                 to_rgb_fcn=None,
                 ignore_index=None,
                 optimizer_name='adam'):
        super().__init__(in_channels,
                         n_classes,
                         lr,
                         log_image_iter,
                         to_rgb_fcn,
                         ignore_index=ignore_index,
                         optimizer_name=optimizer_name)

    def forward(self, batch):
        images = [batch['image']]
        keys = list(batch.keys())
        for k in self.EXTRA_KEYS:
            if k in keys:
                images.append(batch[k])
        return self.model.forward_fused(images)


MODELS = {
    'ms_model': WaterSegmentationModel,
    'ef_model': EarlyFusionModel,
}
# 'lf_model' is registered by lf_model.py (imported at the bottom of this module: it subclasses
# WaterSegmentationModel, so it has to come after the class definitions)


def build_model(model_name, input_channels, n_classes, lr, log_image_iter, to_rgb_fcn, ignore_index,
                **kwargs):
    """Same positional order as the reference factory (models/__init__.py:12-20)."""
    try:
        model = MODELS[model_name](input_channels, n_classes, lr, log_image_iter, to_rgb_fcn,
                                   ignore_index, **kwargs)
    except KeyError:
        print(f'Could not find model named: {model_name}')
        raise
    return model


def install_into_reference() -> None:
    """Make the reference's own scripts (fit.py / infer.py / predict.py) pick up the B200
    classes without editing them: patch ``st_water_seg.models`` after it is imported."""
    import importlib
    ref = importlib.import_module("st_water_seg.models")
    from .unet import UNetDecoder, UNetEncoder
    LateFusionModel = MODELS['lf_model']
    ref.MODELS['ms_model'] = WaterSegmentationModel
    ref.MODELS['ef_model'] = EarlyFusionModel
    ref.MODELS['lf_model'] = LateFusionModel
    for modname, cls in (("st_water_seg.models.unet", UNet),
                         ("st_water_seg.models.unet", UNetEncoder),
                         ("st_water_seg.models.unet", UNetDecoder),
                         ("st_water_seg.models.water_seg_model", WaterSegmentationModel),
                         ("st_water_seg.models.ef_model", EarlyFusionModel),
                         ("st_water_seg.models.lf_model", LateFusionModel)):
        try:
            setattr(importlib.import_module(modname), cls.__name__, cls)
        except Exception:
            pass


from . import lf_model as _lf_model  # noqa: E402,F401  (registers MODELS['lf_model'])
