"""The reference's training entry point, fit.py:16-103, re-enacted from the CALLER's side against the drop-in:
DataLoader of sample dicts -> `build_model(cfg.model.name, n_channels, n_classes, lr, log_image_iter=, to_rgb_fcn=,
ignore_index=)` -> Trainer(max_epochs, ModelCheckpoint(monitor='val_MulticlassJaccardIndex', mode='max')).fit ->
best checkpoint -> infer.py:86-107 (`build_model`, `load_from_checkpoint(path, in_channels=, n_classes=, lr=)`,
`_set_model_to_eval()`, `.to('cuda')`).  pytorch_lightning itself is absent in this image (no network); the loop
is tests/trainer_standin.py."""
import sys
from pathlib import Path

import pytest
import torch
from torch.utils.data import DataLoader, Dataset

sys.path.insert(0, str(Path(__file__).resolve().parent))
from trainer_standin import ModelCheckpoint, Trainer  # noqa: E402

from oracle import unet_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


class SyntheticFloodChips(Dataset):
    """Sample dicts shaped like Floodplanet_Dataset.__getitem__ (floodplanet.py:600-658): 'image' f32 [C,H,W],
    'target' int64 [H,W] (0 = no-flood/nodata, 1 = flood), 'mean' / 'std' [C,1,1]; the water mask is a function
    of the image so that there is something to learn."""
    n_classes = 3
    ignore_index = 0
    n_channels = {"ms_image": 4, "dem": 1}

    def __init__(self, n, size, seed):
        g = torch.Generator().manual_seed(seed)
        coarse = torch.rand(n, 1, size // 8, size // 8, generator=g)
        field = torch.nn.functional.interpolate(coarse, size=(size, size), mode="bilinear", align_corners=False)
        self.target = (field[:, 0] > 0.5).long()
        self.image = torch.rand(n, 4, size, size, generator=g) * 0.5 + field * 0.5
        self.dem = torch.rand(n, 1, size, size, generator=g)

    def __len__(self):
        return self.image.shape[0]

    def __getitem__(self, i):
        return {"image": self.image[i], "dem": self.dem[i], "target": self.target[i],
                "mean": torch.zeros(4, 1, 1), "std": torch.ones(4, 1, 1)}

    @staticmethod
    def to_RGB(image):
        return image[:3]


def test_fit_py_flow_then_infer_py_checkpoint_reload(tmp_path):
    from floodplanet_code_b200.water_seg_model import EarlyFusionModel, build_model
    torch.manual_seed(0)
    train_dataset, valid_dataset = SyntheticFloodChips(40, 64, 1), SyntheticFloodChips(20, 64, 2)
    train_loader = DataLoader(train_dataset, batch_size=10, shuffle=True, num_workers=0)      # conf/config.yaml:21
    valid_loader = DataLoader(valid_dataset, batch_size=10, shuffle=False, num_workers=0)
    model = build_model("ef_model", train_dataset.n_channels, train_dataset.n_classes, 1e-3,          # fit.py:66-73
                        log_image_iter=50, to_rgb_fcn=train_dataset.to_RGB, ignore_index=train_dataset.ignore_index)
    assert isinstance(model, EarlyFusionModel)
    ckpt = ModelCheckpoint(dirpath=str(tmp_path / "checkpoints"), save_top_k=3, mode="max",
                           monitor="val_MulticlassJaccardIndex")
    trainer = Trainer(max_epochs=4, callbacks=[ckpt])
    trainer.fit(model=model, train_dataloaders=train_loader, val_dataloaders=valid_loader)
    losses = trainer.train_losses
    assert len(losses) == 16 and sum(losses[-4:]) < sum(losses[:4])          # it trains through the seam
    assert len(ckpt.saved) == 3 and Path(trainer.checkpoint_callback.best_model_path).exists()
    for key in ("train_MulticlassF1Score", "val_MulticlassJaccardIndex", "val_MulticlassAccuracy", "valid_loss"):
        assert key in model.logged, key
    # ---- infer.py:86-107
    best = trainer.checkpoint_callback.best_model_path
    m2 = build_model("ef_model", train_dataset.n_channels, train_dataset.n_classes, 1e-3, 50, None, 0)
    m2 = m2.load_from_checkpoint(best, in_channels=train_dataset.n_channels, n_classes=3, lr=1e-3)
    m2._set_model_to_eval()
    m2 = m2.to("cuda")
    batch = next(iter(valid_loader))
    batch = {k: v.to("cuda") for k, v in batch.items()}
    with torch.no_grad():
        out2 = m2(batch).detach().cpu().numpy()
    assert out2.dtype.name == "float32" and out2.shape == (10, 3, 64, 64)
    # the reloaded model reproduces the saved one bit for bit (state_dict incl. BatchNorm buffers)
    sd = torch.load(best, weights_only=False)["state_dict"]
    m3 = build_model("ef_model", train_dataset.n_channels, 3, 1e-3, 50, None, 0)
    m3.load_state_dict(sd, strict=True)
    m3._set_model_to_eval()
    with torch.no_grad():
        out3 = m3.to("cuda")(batch)
    assert torch.equal(torch.from_numpy(out2), out3.cpu())
    # and the fp32 oracle on the same checkpoint agrees with the eval forward inside the bf16 envelope
    ref = O.unet_forward({k[len("model."):]: v.cuda() for k, v in sd.items() if k.startswith("model.")},
                         O.early_fusion_input({"image": batch["image"], "dem": batch["dem"]}), training=False)
    rel = float((out3 - ref).norm() / ref.norm())
    assert rel < 8e-2, rel        # same eval-mode envelope as tests/test_unet_gpu.py (bf16 storage through 18 layers)
