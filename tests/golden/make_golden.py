"""Generate the committed golden fixtures from the REFERENCE ITSELF.

Runs only in the build container, where the reference checkout exists at /root/reference:
imports ``st_water_seg/models/unet.py`` and ``st_water_seg/datasets/utils.py`` by file path
(the package import needs pytorch_lightning / hydra, which are not installed), runs them on
CPU fp32 with fixed seeds and stores small input/output vectors in ``tests/golden/*.pt``.
The GPU box has no /root/reference; tests there read only the fixtures.

    python tests/golden/make_golden.py          # everything
    python tests/golden/make_golden.py lf       # only the late-fusion fixture
    python tests/golden/make_golden.py augment  # only the normalise / augment fixture
    python tests/golden/make_golden.py envelope | trajectory | stitch
"""
import importlib.util
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

REF = Path("/root/reference/st_water_seg")
OUT = Path(__file__).resolve().parent


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def summarize_grads(named_grads):
    out = {}
    for k, g in named_grads.items():
        g = g.detach().double().flatten()
        out[k] = {"norm": float(g.norm()), "sum": float(g.sum()), "head": g[:8].float().clone()}
    return out


def unet_case(ref_unet, name, n, c, h, w, n_classes, ignore_index, seed, extra=None, block=8):
    sys.path.insert(0, str(OUT.parent.parent))
    from oracle import unet_oracle as O
    torch.manual_seed(seed)
    model = ref_unet.UNet(c, n_classes)
    init_sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = O.synthetic_batch(n, c, h, w, seed=seed + 1, block=block)
    if extra == "all_ignored":
        batch["target"].zero_()
    model.train()
    logits = model(batch["image"])
    loss = nn.CrossEntropyLoss(ignore_index=ignore_index)(logits, batch["target"])
    if torch.isnan(loss):
        loss = torch.nan_to_num(loss)
    pred = logits.argmax(dim=1)
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
    after = model.state_dict()
    model.eval()
    with torch.no_grad():
        logits_eval = model(batch["image"])
    fx = {
        "cfg": dict(n=n, c=c, h=h, w=w, n_classes=n_classes, ignore_index=ignore_index, seed=seed),
        "init_checksum": {k: float(v.double().sum()) for k, v in init_sd.items()},
        "image": batch["image"], "target": batch["target"],
        "logits_train": logits.detach().clone(), "loss": float(loss.detach()), "pred": pred.clone(),
        "grads": summarize_grads(grads),
        "grad_full": {k: grads[k].detach().clone() for k in
                      ("inc.double_conv.0.weight", "outc.conv.weight", "outc.conv.bias",
                       "inc.double_conv.1.weight", "inc.double_conv.1.bias",
                       "up4.conv.double_conv.4.weight", "up4.conv.double_conv.3.bias")},
        "bn_after": {k: after[k].clone() for k in after if k.startswith("inc.double_conv.1.")
                     or k.startswith("down4.maxpool_conv.1.double_conv.4.")},
        "logits_eval": logits_eval.clone(),
    }
    torch.save(fx, OUT / f"{name}.pt")
    print(name, "loss", float(loss.detach()), "ignored_frac", float((batch["target"] == 0).float().mean()), "bytes", (OUT / f"{name}.pt").stat().st_size)


def op_semantics_case():
    """Edge semantics of the third-party ops the path relies on."""
    import torch.nn.functional as F
    fx = {}
    z = torch.zeros(1, 1, 4, 4)
    _, idx = F.max_pool2d(z, 2, return_indices=True)
    fx["maxpool_zero_idx"] = idx.flatten().tolist()
    zn = z.clone(); zn[0, 0, 1, 1] = float("nan")
    p, idx = F.max_pool2d(zn, 2, return_indices=True)
    fx["maxpool_nan_idx"] = idx.flatten().tolist()
    fx["maxpool_nan_isnan"] = torch.isnan(p).flatten().tolist()
    lg = torch.zeros(1, 3, 2, 2)
    lg[0, :, 0, 0] = torch.tensor([1.0, float("nan"), 2.0])
    fx["argmax_ties"] = lg.argmax(1).flatten().tolist()
    t0 = torch.zeros(1, 2, 2, dtype=torch.long)
    lgr = torch.randn(1, 3, 2, 2, requires_grad=True)
    l = F.cross_entropy(lgr, t0, ignore_index=0)
    fx["ce_all_ignored_isnan"] = bool(torch.isnan(l))
    torch.nan_to_num(l).backward()
    fx["ce_all_ignored_grad_abs_sum"] = float(lgr.grad.abs().sum())
    ramp = torch.arange(4.0).view(1, 1, 1, 4)
    fx["bilinear_4_to_8"] = F.interpolate(ramp, scale_factor=(1, 2), mode="bilinear", align_corners=True).flatten()
    x = torch.randn(4, 2, 3, 3, generator=torch.Generator().manual_seed(0))
    rm, rv = torch.zeros(2), torch.ones(2)
    F.batch_norm(x, rm, rv, None, None, True, 0.1, 1e-5)
    fx["bn_x"] = x
    fx["bn_running_mean"] = rm
    fx["bn_running_var"] = rv
    torch.save(fx, OUT / "op_semantics.pt")
    print("op_semantics", fx["maxpool_zero_idx"], fx["argmax_ties"])


def tiler_case():
    hydra = types.ModuleType("hydra"); hydra.utils = types.ModuleType("hydra.utils")
    hydra.utils.get_original_cwd = lambda: "."
    sys.modules.setdefault("hydra", hydra); sys.modules.setdefault("hydra.utils", hydra.utils)
    ref_utils = load_by_path("ref_ds_utils", REF / "datasets" / "utils.py")
    cases = [(10240, 10240, 512, 512, 512), (10240, 10240, 512, 512, 256), (1024, 1024, 300, 300, 150),
             (700, 530, 300, 300, 300), (1000, 900, 512, 512, 512)]
    fx = {}
    for (H, W, ch, cw, st) in cases:
        sl = ref_utils.get_crop_slices(H, W, ch, cw, st, mode="exact")
        fx[(H, W, ch, cw, st)] = {"count": len(sl), "head": sl[:5], "tail": sl[-5:],
                                  "checksum": int(sum((i + 1) * (a + 3 * b + 5 * c + 7 * d)
                                                      for i, (a, b, c, d) in enumerate(sl)))}
    torch.save(fx, OUT / "tiler.pt")
    print("tiler", {k: v["count"] for k, v in fx.items()})


def _stub_reference_deps():
    """The reference package imports pytorch_lightning / torchmetrics / omegaconf-based tools at
    module import; none is installed here and none takes part in the arithmetic of forward /
    loss / backward, so minimal stand-ins let the UNMODIFIED st_water_seg.models.lf_model (and
    water_seg_model, unet) be imported and run."""
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = nn.Module
    sys.modules.setdefault("pytorch_lightning", pl)

    tm = types.ModuleType("torchmetrics")

    class _Metric:
        def __init__(self, *a, **k):
            pass

    class _Collection(_Metric):
        def clone(self, prefix=None):
            return self

    tm.MetricCollection = _Collection
    tm.F1Score = tm.JaccardIndex = tm.Accuracy = _Metric
    sys.modules.setdefault("torchmetrics", tm)
    tools = types.ModuleType("st_water_seg.tools")
    tools.create_conf_matrix_pred_image = lambda *a, **k: None
    sys.modules.setdefault("st_water_seg.tools", tools)
    if str(REF.parent) not in sys.path:
        sys.path.insert(0, str(REF.parent))


def lf_case(name, in_channels, n, h, w, n_classes, ignore_index, seed, block=8):
    """Late fusion: the reference's own LateFusionModel (lf_model.py) on CPU fp32."""
    _stub_reference_deps()
    from st_water_seg.models.lf_model import LateFusionModel
    sys.path.insert(0, str(OUT.parent.parent))
    from oracle import unet_oracle as O
    torch.manual_seed(seed)
    model = LateFusionModel(dict(in_channels), n_classes, 1e-4, ignore_index=ignore_index)
    init_sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = O.synthetic_batch(n, in_channels["ms_image"], h, w, seed=seed + 1, block=block)
    g = torch.Generator().manual_seed(seed + 2)
    for key, c in in_channels.items():
        if key != "ms_image":
            batch[key] = torch.rand(n, c, h, w, generator=g)
    model._set_model_to_train()
    logits = model.forward(batch)
    loss = model.loss_func(logits, batch["target"])
    if torch.isnan(loss):
        loss = torch.nan_to_num(loss)
    pred = logits.argmax(dim=1)
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
    after = model.state_dict()
    model._set_model_to_eval()
    with torch.no_grad():
        logits_eval = model.forward(batch)
        enc_feats = model.encoders["ms_image"](batch["image"])
    fx = {
        "cfg": dict(in_channels=dict(in_channels), n=n, h=h, w=w, n_classes=n_classes,
                    ignore_index=ignore_index, seed=seed),
        "state_dict_keys": list(init_sd.keys()),
        "init_checksum": {k: float(v.double().sum()) for k, v in init_sd.items()},
        "batch": batch,
        "logits_train": logits.detach().clone(), "loss": float(loss.detach()), "pred": pred.clone(),
        "grads": summarize_grads(grads),
        "grad_full": {k: grads[k].detach().clone() for k in
                      ("concat_convs.0.weight", "concat_convs.0.bias", "concat_convs.4.bias",
                       "decoder.outc.conv.weight", "encoders.dem.inc.double_conv.0.weight",
                       "encoders.ms_image.inc.double_conv.1.weight")},
        "bn_after": {k: after[k].clone() for k in after if k.startswith("encoders.dem.inc.double_conv.1.")},
        "logits_eval": logits_eval.clone(),
        "enc_feat_eval_checksum": [float(f.double().sum()) for f in enc_feats],
        "enc_feat_eval_x5": enc_feats[4].clone(),
    }
    torch.save(fx, OUT / f"{name}.pt")
    print(name, "loss", float(loss.detach()), "bytes", (OUT / f"{name}.pt").stat().st_size)


def augment_case():
    """Outputs of the reference's OWN `BaseDataset.normalize / sample_transforms / apply_transforms`
    (datasets/base_dataset.py:77-113,494-555), the class loaded by file path with stand-ins for the
    modules that are not installed (tifffile, pytorch_lightning) or that only do file IO
    (st_water_seg.datasets.utils); none of them touches the arithmetic under test."""
    import numpy as np
    tf = types.ModuleType("tifffile"); tf.tifffile = types.ModuleType("tifffile.tifffile")
    sys.modules.setdefault("tifffile", tf); sys.modules.setdefault("tifffile.tifffile", tf.tifffile)
    pl = types.ModuleType("pytorch_lightning"); pl.seed_everything = lambda s: None
    sys.modules.setdefault("pytorch_lightning", pl)
    for name in ("st_water_seg", "st_water_seg.datasets"):
        sys.modules.setdefault(name, types.ModuleType(name))
    du = types.ModuleType("st_water_seg.datasets.utils")
    du.load_global_dataset_norm_params = lambda name: None
    sys.modules["st_water_seg.datasets.utils"] = du
    bd = load_by_path("ref_base_dataset", REF / "datasets" / "base_dataset.py")

    class Cfg(dict):                         # attribute access like the OmegaConf node the reference reads
        __getattr__ = dict.__getitem__

    def cfg(h, v, r, lo=0, hi=360):
        return Cfg(hflip=Cfg(active=True, likelihood=h), vflip=Cfg(active=True, likelihood=v),
                   rotate=Cfg(active=True, likelihood=r, min_rot_angle=lo, max_rot_angle=hi))

    cases = []
    rng = np.random.RandomState(123)
    specs = [  # (C, H, W, norm_mode, hflip p, vflip p, rotate p, numpy seed)
        (4, 64, 64, None, 1.0, 0.0, 0.0, 1), (4, 64, 64, None, 0.0, 1.0, 0.0, 2),
        (4, 64, 64, None, 0.0, 0.0, 1.0, 3), (4, 44, 36, "local", 1.0, 1.0, 1.0, 4),
        (6, 75, 75, "global", 0.5, 0.5, 0.5, 5), (4, 75, 75, "local", 0.5, 0.5, 0.5, 6),
        (2, 128, 128, None, 0.5, 0.5, 1.0, 7), (2, 37, 53, "global", 0.0, 1.0, 1.0, 8),
        (1, 300, 300, None, 1.0, 0.0, 1.0, 9), (4, 64, 64, None, 0.0, 0.0, 0.0, 10),
    ]
    for (c, h, w, norm_mode, ph, pv, pr, seed) in specs:
        ds = bd.BaseDataset.__new__(bd.BaseDataset)     # the methods under test read only these attributes
        ds.norm_mode = norm_mode
        ds.transforms = cfg(ph, pv, pr)
        gp = None
        if norm_mode == "global":
            gp = {"mean": rng.rand(c), "std": rng.rand(c) + 0.5}       # float64, as np.mean / np.std of the sampled pixels give
            ds.global_norm_params = {"PS": gp}
        image = rng.rand(c, h, w).astype(np.float32)
        target = (rng.rand(h, w) < 0.4).astype(np.int64)
        np.random.seed(seed)
        norm_image, mean, std = ds.normalize(image.copy(), "PS")
        active = ds.sample_transforms()
        out_img = ds.apply_transforms(norm_image, active, is_anno=False).float()
        out_tgt = ds.apply_transforms(target, active, is_anno=True).long()
        cases.append({
            "image": torch.from_numpy(image), "target": torch.from_numpy(target).to(torch.int8), "norm_mode": norm_mode,
            "global_params": None if gp is None else {k: torch.from_numpy(v) for k, v in gp.items()},
            "cfg": {k: dict(v) for k, v in ds.transforms.items()}, "np_seed": seed,
            "active": [{"transform": t["transform"].__name__, "kwargs": {k: float(v) for k, v in t["kwargs"].items()}}
                       for t in active],
            "mean": torch.as_tensor(np.asarray(mean)).flatten().double(),
            "std": torch.as_tensor(np.asarray(std)).flatten().double(),
            "out_image": out_img, "out_target": out_tgt.to(torch.int8),   # labels stored as int8 to keep the file small
        })
        print(f"augment case C{c} {h}x{w} norm={norm_mode}: active={[t['transform'].__name__ for t in active]}")
    torch.save({"cases": cases, "torchvision": __import__("torchvision").__version__}, OUT / "augment.pt")
    print(f"wrote {OUT / 'augment.pt'}")


def stitch_case():
    """Outputs of the reference's OWN `ImageStitcher_v2.add_image / get_combined_images`
    (utils/utils_image.py:364-494,569-571) and of infer.py's mask rule (:181-184) for tile sets made by
    the reference's own `get_crop_slices` + `CropParams` (datasets/utils.py:22-52,86-212): a stride = crop
    case, a ragged one and an overlapping one.  `utils_image.py` is loaded by file path with a stand-in
    for `tifffile` (only used by save_images, not by the arithmetic under test)."""
    import tempfile
    import numpy as np
    tf = types.ModuleType("tifffile"); tf.tifffile = types.ModuleType("tifffile.tifffile")
    sys.modules.setdefault("tifffile", tf); sys.modules.setdefault("tifffile.tifffile", tf.tifffile)
    hydra = types.ModuleType("hydra"); hydra.utils = types.ModuleType("hydra.utils")
    hydra.utils.get_original_cwd = lambda: "."
    sys.modules.setdefault("hydra", hydra); sys.modules.setdefault("hydra.utils", hydra.utils)
    ui = load_by_path("ref_utils_image", REF / "utils" / "utils_image.py")
    du = load_by_path("ref_ds_utils2", REF / "datasets" / "utils.py")
    cases = []
    for (H, W, crop, stride, ncls, seed) in [(96, 64, 32, 32, 3, 1), (100, 70, 32, 32, 3, 2), (80, 112, 32, 16, 3, 3),
                                              (75, 75, 30, 15, 2, 4)]:
        tiles = du.get_crop_slices(H, W, crop, crop, stride, mode="exact")
        rng = np.random.RandomState(seed)
        with tempfile.TemporaryDirectory() as td:
            st = ui.ImageStitcher_v2(td, save_backend="tifffile", save_ext=".tif")
            preds = []
            for (h0, w0, hh, ww) in tiles:
                # what infer.py:122-127 hands the stitcher: softmax over classes of a full crop-sized output
                lg = rng.standard_normal((ncls, crop, crop)).astype(np.float32) * 2
                from scipy.special import softmax
                pred = softmax(lg[None], axis=1)[0].transpose(1, 2, 0)          # infer.py:123,127
                preds.append(lg)
                cp = du.CropParams(h0, w0, hh, ww, H, W, crop, crop)
                st.add_image(pred, "scene", cp, H, W)
            canvas = st.get_combined_images()["scene"]
        mask = (np.clip(canvas.argmax(axis=2), 0, 1) * 255).astype("uint8")      # infer.py:181-184
        cases.append({"H": H, "W": W, "crop": crop, "stride": stride, "n_classes": ncls, "seed": seed,
                      "tiles": [list(map(int, t)) for t in tiles],
                      # logits are regenerated in the test: RandomState(seed).standard_normal((ncls, crop, crop))
                      # .astype(float32) * 2 per tile, in tile order
                      "logits_checksum": float(np.stack(preds).astype(np.float64).sum()),
                      "canvas": torch.from_numpy(np.asarray(canvas)), "mask": torch.from_numpy(mask)})
        print(f"stitch case {H}x{W} crop {crop} stride {stride}: {len(tiles)} tiles, canvas {canvas.dtype}, "
              f"water fraction {(mask > 0).mean():.3f}")
    torch.save({"cases": cases}, OUT / "stitch.pt")
    print("wrote", OUT / "stitch.pt", (OUT / "stitch.pt").stat().st_size, "bytes")


def _rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _ref_step(ref_unet, init_sd, batch, c, n_classes, ignore_index, autocast):
    """One forward + loss + backward of the UNMODIFIED reference module, fp32 or under stock
    torch.autocast('cpu', bfloat16)."""
    model = ref_unet.UNet(c, n_classes)
    model.load_state_dict(init_sd)
    model.train()
    if autocast:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            logits = model(batch["image"])
            loss = nn.CrossEntropyLoss(ignore_index=ignore_index)(logits.float(), batch["target"])
    else:
        logits = model(batch["image"])
        loss = nn.CrossEntropyLoss(ignore_index=ignore_index)(logits, batch["target"])
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    return logits.detach().float(), float(loss.detach()), grads


def envelope_case(only=None):
    """Third point of the whole-network comparison: how far is ANOTHER bf16 implementation -- the
    unmodified reference module under stock torch.autocast(bfloat16) -- from the reference's own
    fp32 result, on the same weights and inputs?  Stored as relative L2 errors (logits, loss, every
    parameter gradient) in a small JSON; tests/test_envelope_gpu.py asserts that the CUDA path is no
    further from fp32 than 1.25x this envelope.  Inputs / weights are regenerated from seeds
    (oracle.synthetic_batch / init_state_dict == torch.manual_seed(s); UNet(...))."""
    import json
    sys.path.insert(0, str(OUT.parent.parent))
    from oracle import unet_oracle as O
    ref_unet = load_by_path("ref_unet", REF / "models" / "unet.py")
    # the first three are the inputs / weights of the committed unet_* fixtures above
    cases = [dict(name="unet_c4_32", n=2, c=4, h=32, w=32, seed=0, block=8),
             dict(name="unet_c4_44x36", n=1, c=4, h=44, w=36, seed=3, block=8),
             dict(name="unet_c6_37_ef", n=2, c=6, h=37, w=37, seed=5, block=8, n_classes=2, ignore_index=-100),
             dict(name="b2_128", n=2, c=4, h=128, w=128, seed=20, block=16),
             dict(name="b4_256", n=4, c=4, h=256, w=256, seed=21, block=32),
             dict(name="b2_300", n=2, c=4, h=300, w=300, seed=22, block=20),
             dict(name="parity_b8_512", n=8, c=4, h=512, w=512, seed=40, block=32)]   # BASELINE.json configs[0]
    out = {"how": "reference UNet (st_water_seg/models/unet.py) + CrossEntropyLoss(ignore_index=0), CPU; "
                  "rel = ||autocast - fp32|| / ||fp32||; torch " + torch.__version__,
           "cases": []}
    path = OUT / "autocast_envelope.json"
    if only:                                   # regenerate some cases, keep the others
        out = json.loads(path.read_text())
        out["cases"] = [c for c in out["cases"] if c["name"] not in only]
        cases = [c for c in cases if c["name"] in only]
    for cs in cases:
        ncls, ii = cs.setdefault("n_classes", 3), cs.setdefault("ignore_index", 0)
        torch.manual_seed(cs["seed"])
        init_sd = {k: v.detach().clone() for k, v in ref_unet.UNet(cs["c"], ncls).state_dict().items()}
        chk = O.init_state_dict(cs["c"], ncls, seed=cs["seed"])
        assert all(torch.equal(init_sd[k], chk[k]) for k in init_sd)       # the seeds regenerate the weights
        batch = O.synthetic_batch(cs["n"], cs["c"], cs["h"], cs["w"], seed=cs["seed"] + 1, block=cs["block"])
        lg32, loss32, g32 = _ref_step(ref_unet, init_sd, batch, cs["c"], ncls, ii, autocast=False)
        lg16, loss16, g16 = _ref_step(ref_unet, init_sd, batch, cs["c"], ncls, ii, autocast=True)
        rec = dict(cs)
        rec["loss_fp32"] = loss32
        rec["loss_autocast"] = loss16
        rec["logits_rel"] = _rel(lg16, lg32)
        rec["argmax_agree"] = float((lg16.argmax(1) == lg32.argmax(1)).float().mean())
        rec["grad_rel"] = {k: _rel(g16[k], g32[k]) for k in g32}
        rec["grad_cos"] = {k: float((g16[k].double().flatten() @ g32[k].double().flatten())
                                    / (g16[k].double().norm() * g32[k].double().norm()).clamp_min(1e-30)) for k in g32}
        out["cases"].append(rec)
        w = [k for k in g32 if k.endswith("weight") and g32[k].dim() == 4]
        print(f"envelope {cs}: logits {rec['logits_rel']:.4f} loss {loss32:.5f}/{loss16:.5f} "
              f"grad rel min {min(rec['grad_rel'][k] for k in w):.3f} max {max(rec['grad_rel'][k] for k in w):.3f}")
    path.write_text(json.dumps(out, indent=1))


def trajectory_case(steps=100, n=8, size=64, lr=1e-3, seed=30):
    """Training trajectory of the UNMODIFIED reference module: Adam (water_seg_model.py:198-205) on a
    fixed set of 8 chips for 100 steps, in fp32 and under stock torch.autocast(bfloat16); loss per step,
    final eval-mode confusion counts.  tests/test_trajectory_gpu.py trains the CUDA path on the same
    chips from the same weights and compares the curves."""
    import json
    sys.path.insert(0, str(OUT.parent.parent))
    from oracle import unet_oracle as O
    ref_unet = load_by_path("ref_unet", REF / "models" / "unet.py")
    torch.manual_seed(seed)
    init_sd = {k: v.detach().clone() for k, v in ref_unet.UNet(4, 3).state_dict().items()}
    batch = O.synthetic_batch(n, 4, size, size, seed=seed + 1, block=8, ignore_frac=0.5)
    out = {"cfg": dict(steps=steps, n=n, size=size, lr=lr, seed=seed, ignore_index=-100, ignore_frac=0.5, block=8),
           "how": "reference UNet + CrossEntropyLoss (no pixel ignored: classes 0 / 1) + optim.Adam(lr), CPU, torch "
                  + torch.__version__}
    for mode in ("fp32", "autocast"):
        torch.manual_seed(0)
        model = ref_unet.UNet(4, 3)
        model.load_state_dict(init_sd)
        opt = torch.optim.Adam(model.parameters(), lr=lr)
        lossf = nn.CrossEntropyLoss(ignore_index=-100)
        curve = []
        for i in range(steps):
            model.train()
            opt.zero_grad()
            if mode == "autocast":
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    loss = lossf(model(batch["image"]).float(), batch["target"])
            else:
                loss = lossf(model(batch["image"]), batch["target"])
            loss.backward()
            opt.step()
            curve.append(float(loss.detach()))
        model.eval()
        with torch.no_grad():
            if mode == "autocast":
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    lg = model(batch["image"]).float()
            else:
                lg = model(batch["image"])
        conf = O.confusion_counts(lg.argmax(1).flatten(), batch["target"].flatten(), 3, None)
        mm = O.micro_metrics(conf, None)
        out[mode] = {"loss": curve, "eval_loss": float(lossf(lg, batch["target"])), "confusion": conf.tolist(),
                     "metrics": mm}
        print(f"trajectory {mode}: loss {curve[0]:.4f} -> {curve[-1]:.4f}, eval loss {out[mode]['eval_loss']:.4f}, {mm}")
    (OUT / "trajectory.json").write_text(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "envelope":
        envelope_case(sys.argv[2:] or None)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "stitch":
        stitch_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "trajectory":
        trajectory_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "lf":
        lf_case("lf_c4_dem1_32", {"ms_image": 4, "dem": 1}, n=2, h=32, w=32, n_classes=3, ignore_index=0, seed=11)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "augment":
        augment_case()
        sys.exit(0)
    ref_unet = load_by_path("ref_unet", REF / "models" / "unet.py")
    unet_case(ref_unet, "unet_c4_32", n=2, c=4, h=32, w=32, n_classes=3, ignore_index=0, seed=0)
    unet_case(ref_unet, "unet_c4_44x36", n=1, c=4, h=44, w=36, n_classes=3, ignore_index=0, seed=3)
    unet_case(ref_unet, "unet_c6_37_ef", n=2, c=6, h=37, w=37, n_classes=2, ignore_index=-100, seed=5)
    unet_case(ref_unet, "unet_c4_32_allignored", n=1, c=4, h=32, w=32, n_classes=3, ignore_index=0, seed=7,
              extra="all_ignored")
    op_semantics_case()
    tiler_case()
    lf_case("lf_c4_dem1_32", {"ms_image": 4, "dem": 1}, n=2, h=32, w=32, n_classes=3, ignore_index=0, seed=11)
    augment_case()
    envelope_case()
    trajectory_case()
    stitch_case()
