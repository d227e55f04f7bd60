"""The drop-in seam against the REAL reference package (CPU; runs only where /root/reference exists,
i.e. in the build container -- the GPU box has no reference checkout).

`install_into_reference()` must make the reference's own, unmodified `st_water_seg.models.build_model`
(models/__init__.py:12-20, called by fit.py:66-73 / infer.py:86-93 / predict.py:164-171) return the B200
classes for every registered model name, with the reference's positional call, and a `state_dict` saved
by the REFERENCE classes must load into them with strict=True (checkpoint compatibility,
infer.py:96-99).  Runs in a subprocess: the reference needs import-time stand-ins for
pytorch_lightning / torchmetrics (make_golden._stub_reference_deps) that must not leak into this
process."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/st_water_seg")

SCRIPT = r'''
import importlib.util, sys, io
sys.path.insert(0, "{root}")
import torch
# the product first: it must bind its OWN Lightning stand-in, not the stub the reference gets below
import floodplanet_code_b200.water_seg_model as b200
from floodplanet_code_b200.lf_model import LateFusionModel
from floodplanet_code_b200.unet import UNet, UNetEncoder, UNetDecoder
spec = importlib.util.spec_from_file_location("make_golden", "{root}/tests/golden/make_golden.py")
mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
mg._stub_reference_deps()
import st_water_seg.models as ref                       # the unmodified reference package
import st_water_seg.models.unet as ref_unet
assert ref.__file__.startswith("/root/reference/")
cases = [("ms_model", {{"ms_image": 4}}), ("ef_model", {{"ms_image": 4, "dem": 1, "slope": 1}}),
         ("lf_model", {{"ms_image": 4, "dem": 1}})]
saved = {{}}
for name, ch in cases:                                   # 1. the reference's own classes, before the switch
    torch.manual_seed(3)
    m = ref.build_model(name, ch, 3, 1e-4, 50, None, 0)
    assert type(m).__module__.startswith("st_water_seg."), type(m)
    buf = io.BytesIO(); torch.save({{"state_dict": m.state_dict()}}, buf); saved[name] = buf.getvalue()
ref_classes = {{n: ref.MODELS[n] for n, _ in cases}}
b200.install_into_reference()                            # 2. the one-line switch of INTEGRATION.md
want = {{"ms_model": b200.WaterSegmentationModel, "ef_model": b200.EarlyFusionModel, "lf_model": LateFusionModel}}
for name, ch in cases:
    m = ref.build_model(name, ch, 3, 1e-4, 50, None, 0)   # the REFERENCE's factory, reference call order
    assert type(m) is want[name] and type(m) is not ref_classes[name], (name, type(m))
    assert m.ignore_index == 0 and m.n_classes == 3 and m.lr == 1e-4 and m.log_image_iter == 50
    sd = torch.load(io.BytesIO(saved[name]), weights_only=False)["state_dict"]
    missing = m.load_state_dict(sd, strict=True)         # checkpoint of the reference loads unchanged
    own = m.state_dict()
    assert list(own.keys()) == list(sd.keys()), name
    assert all(own[k].dtype == sd[k].dtype and own[k].shape == sd[k].shape and torch.equal(own[k], sd[k]) for k in sd)
    assert isinstance(m.configure_optimizers(), torch.optim.Adam)
    for attr in ("model" if name != "lf_model" else "decoder", "loss_func", "train_metrics", "valid_metrics",
                 "test_metrics", "training_step", "validation_step", "test_step", "_set_model_to_train",
                 "_set_model_to_eval", "load_from_checkpoint"):
        assert hasattr(m, attr), (name, attr)
m = ref.build_model("ms_model", {{"ms_image": 4}}, 3, 1e-4, 50, None, -1)
assert m.ignore_index == 2                                # water_seg_model.py:35-36
assert ref_unet.UNet is UNet and ref_unet.UNetEncoder is UNetEncoder and ref_unet.UNetDecoder is UNetDecoder
import st_water_seg.models.water_seg_model as rw, st_water_seg.models.ef_model as re_, st_water_seg.models.lf_model as rl
assert rw.WaterSegmentationModel is b200.WaterSegmentationModel and re_.EarlyFusionModel is b200.EarlyFusionModel
assert rl.LateFusionModel is LateFusionModel
try:
    ref.build_model("no_such_model", {{"ms_image": 4}}, 3, 1e-4, 50, None, 0)
    raise SystemExit("unknown model name did not raise")
except (KeyError, UnboundLocalError):
    pass      # the reference's factory prints "Could not find model named" and then trips over its unbound
              # local (models/__init__.py:16-20); the product's own build_model re-raises the KeyError
# no CUDA here: the patched classes must fail loudly on a CPU tensor, never fall back
try:
    m({{"image": torch.zeros(1, 4, 32, 32)}})
    raise SystemExit("CPU input did not raise")
except RuntimeError as e:
    assert "no CPU fallback" in str(e)
print("SEAM-OK")
'''


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present (GPU box)")
def test_install_into_reference_build_model_and_checkpoints():
    out = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT)], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and "SEAM-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
