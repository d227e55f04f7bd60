"""A minimal stand-in for the part of pytorch_lightning 1.x that the reference's fit.py drives
(fit.py:76-97: `pl.Trainer(max_epochs, accelerator='gpu', devices=1, callbacks=[ModelCheckpoint(monitor=
'val_MulticlassJaccardIndex', mode='max', save_top_k)]).fit(model, train_dataloaders, val_dataloaders)`).
pytorch_lightning is not installable here (no network), so the tests use this loop to exercise the
LightningModule protocol of the drop-in from the CALLER's side: hook order and arguments follow PL 1.7
(training_step -> backward -> optimizer.step; validation under no_grad with validation_step per batch, then
validation_epoch_end(outputs); checkpoints are {'state_dict': ...} files as `load_from_checkpoint` reads)."""
import os

import torch


class ModelCheckpoint:
    def __init__(self, dirpath, save_top_k=3, mode="max", monitor="val_MulticlassJaccardIndex"):
        self.dirpath, self.save_top_k, self.mode, self.monitor = dirpath, save_top_k, mode, monitor
        self.saved = []                      # (score, path)
        self.best_model_path = ""

    def on_validation_end(self, epoch, module):
        score = float(module.logged[self.monitor])
        os.makedirs(self.dirpath, exist_ok=True)
        path = os.path.join(self.dirpath, f"model-epoch={epoch:02d}-{self.monitor}={score:.4f}.ckpt")
        torch.save({"state_dict": module.state_dict(), "epoch": epoch}, path)
        self.saved.append((score, path))
        self.saved.sort(key=lambda t: -t[0] if self.mode == "max" else t[0])
        for _, p in self.saved[self.save_top_k:]:
            if os.path.exists(p):
                os.remove(p)
        self.saved = self.saved[:self.save_top_k]
        self.best_model_path = self.saved[0][1]


class Trainer:
    def __init__(self, max_epochs, callbacks=(), device="cuda", limit_train_batches=None, limit_val_batches=None):
        self.max_epochs, self.callbacks, self.device = max_epochs, list(callbacks), device
        self.limit_train_batches, self.limit_val_batches = limit_train_batches, limit_val_batches
        self.checkpoint_callback = next((c for c in self.callbacks if isinstance(c, ModelCheckpoint)), None)
        self.train_losses = []

    def _to_device(self, batch):
        return {k: (v.to(self.device) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}

    def fit(self, model, train_dataloaders, val_dataloaders):
        model.to(self.device)
        optimizer = model.configure_optimizers()
        for epoch in range(self.max_epochs):
            model.current_epoch = epoch
            model.train()
            for i, batch in enumerate(train_dataloaders):
                if self.limit_train_batches is not None and i >= self.limit_train_batches:
                    break
                optimizer.zero_grad()
                loss = model.training_step(self._to_device(batch), i)
                loss.backward()
                optimizer.step()
                model.global_step += 1
                self.train_losses.append(float(loss))
            outputs = []
            with torch.no_grad():
                for i, batch in enumerate(val_dataloaders):
                    if self.limit_val_batches is not None and i >= self.limit_val_batches:
                        break
                    outputs.append(model.validation_step(self._to_device(batch), i))
            model.validation_epoch_end(outputs)
            for cb in self.callbacks:
                cb.on_validation_end(epoch, model)
            model.valid_metrics.reset()
            model.train_metrics.reset()
