"""CPU tests (no GPU): the drop-in boundary -- C-ABI symbols, state_dict layout, constructor
signatures, error behaviour -- without launching any kernel."""
import ctypes
import inspect
import re
from pathlib import Path

import pytest
import torch

from oracle import unet_oracle as O

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


def header_symbols():
    text = (ROOT / "include" / "floodplanet_b200.h").read_text()
    return sorted(set(re.findall(r"^(?:int|long)\s+(fpb200_\w+)\s*\(", text, flags=re.M)))


def test_library_builds_loads_and_exports_every_header_symbol():
    from floodplanet_code_b200 import build, capi
    path = build.build_library()
    assert path.exists()
    lib = ctypes.CDLL(str(path))
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in capi.SIGNATURES, f"{s} has no ctypes prototype"
    assert sorted(capi.SIGNATURES) == syms
    assert capi.load().fpb200_abi_version() == 1


def test_kernels_are_blackwell_native_sass():
    """tcgen05 / TMA must be what the conv kernels compile to (UTCHMMA, UTMALDG, LDTM)."""
    import shutil
    import subprocess
    from floodplanet_code_b200 import build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(build.build_library())], capture_output=True,
                          text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass  # no legacy mma.sync path
    # the 64-output-channel halo kernel issues weight-stationary MMAs with collector re-use of the weight slice,
    # the head backward packed fp32 arithmetic and cp.async staging
    for mnemonic in ("UTCHMMA.WS", "B_KEEP", "B_REUSE", "FFMA2", "LDGSTS"):
        assert mnemonic in sass, mnemonic


def test_state_dict_layout_equals_reference():
    from floodplanet_code_b200.unet import UNet
    fx = torch.load(GOLDEN / "unet_c4_32.pt", weights_only=False)
    m = UNet(4, 3)
    sd = m.state_dict()
    assert list(sd.keys()) == list(fx["init_checksum"].keys())  # 128 entries, reference order
    assert len(sd) == 128
    assert sum(p.numel() for p in m.parameters()) == 17268099
    assert sd["inc.double_conv.1.num_batches_tracked"].dtype == torch.int64
    assert sd["inc.double_conv.0.weight"].shape == (64, 4, 3, 3) and sd["inc.double_conv.0.weight"].dtype == torch.float32
    # same default initialisation under the same seed as the reference constructor
    torch.manual_seed(0)
    m0 = UNet(4, 3)
    for k, v in fx["init_checksum"].items():
        assert float(m0.state_dict()[k].double().sum()) == pytest.approx(v, rel=1e-12, abs=1e-12), k
    # oracle weights load strictly
    m.load_state_dict(O.init_state_dict(4, 3, seed=0), strict=True)


def test_constructor_signatures_match_reference():
    from floodplanet_code_b200.unet import UNet
    from floodplanet_code_b200.water_seg_model import (EarlyFusionModel, MODELS,
                                                       WaterSegmentationModel, build_model)
    assert list(inspect.signature(UNet.__init__).parameters) == ["self", "n_channels", "n_classes", "bilinear"]
    want = ["self", "in_channels", "n_classes", "lr", "log_image_iter", "to_rgb_fcn", "ignore_index",
            "optimizer_name"]
    assert list(inspect.signature(WaterSegmentationModel.__init__).parameters) == want
    assert list(inspect.signature(EarlyFusionModel.__init__).parameters) == want
    assert list(inspect.signature(build_model).parameters) == [
        "model_name", "input_channels", "n_classes", "lr", "log_image_iter", "to_rgb_fcn", "ignore_index",
        "kwargs"]
    assert set(MODELS) >= {"ms_model", "ef_model"}
    m = WaterSegmentationModel({"ms_image": 4, "dem": 1}, 3, 1e-4, ignore_index=-1)
    assert m.ignore_index == 2 and m.model.n_channels == 5          # reference :35-36, :79-85
    assert all(k.startswith("model.") for k in m.state_dict())
    for attr in ("loss_func", "train_metrics", "valid_metrics", "test_metrics", "training_step",
                 "validation_step", "test_step", "configure_optimizers", "_set_model_to_train",
                 "_set_model_to_eval", "load_from_checkpoint"):
        assert hasattr(m, attr), attr
    with pytest.raises(NotImplementedError):
        WaterSegmentationModel({"a": 4}, 3, 1e-4, optimizer_name="sgd").configure_optimizers()


def test_no_cpu_fallback():
    from floodplanet_code_b200.loss import MaskedCrossEntropyLoss
    from floodplanet_code_b200.unet import UNet
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        UNet(4, 3)(torch.zeros(1, 4, 32, 32))
    with pytest.raises(RuntimeError):
        MaskedCrossEntropyLoss(0)(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError):
        UNet(4, 3).inc(torch.zeros(1, 4, 32, 32))  # blocks are containers, not eager modules


def test_engine_schedule_matches_survey_flops():
    from floodplanet_code_b200.engine import pad_channels, unet_conv_specs
    specs = unet_conv_specs(4)
    assert len(specs) == 18
    res = {0: 512 * 512, 1: 256 * 256, 2: 128 * 128, 3: 64 * 64, 4: 32 * 32}
    fwd = sum(2 * res[s.level] * s.cout * 9 * s.cin for s in specs) + 2 * res[0] * 3 * 64
    assert fwd / 1e9 == pytest.approx(320.210, rel=1e-4)            # SURVEY.md 2b total
    assert [pad_channels(c) for c in (4, 6, 16, 17, 21, 33)] == [16, 16, 16, 32, 32, 64]


def test_metrics_micro():
    from floodplanet_code_b200.metrics import MicroSegmentationMetrics
    m = MicroSegmentationMetrics(3, ignore_index=0, prefix="train_")
    pred = torch.tensor([1, 1, 2, 0, 1, 2])
    tgt = torch.tensor([1, 0, 1, 1, 1, 0])
    out = m(pred, tgt)
    ref = O.micro_metrics(O.confusion_counts(pred, tgt, 3, 0), ignore_index=0)
    assert float(out["train_MulticlassAccuracy"]) == pytest.approx(ref["Accuracy"])
    assert float(out["train_MulticlassJaccardIndex"]) == pytest.approx(ref["Jaccard"])
    assert float(m.compute()["train_MulticlassF1Score"]) == pytest.approx(ref["F1"])
    # one valid pixel is predicted as the ignored class 0: tp 2, total 4 -> 2 / (8 - 2 - 1)
    assert float(out["train_MulticlassJaccardIndex"]) == pytest.approx(2 / 5)
    # ignore_index outside [0, C) (e.g. None / -100): plain tp / (2 total - tp)
    m2 = MicroSegmentationMetrics(3, ignore_index=None)
    o2 = m2(pred, tgt)
    r2 = O.micro_metrics(O.confusion_counts(pred, tgt, 3, None), None)
    assert float(o2["MulticlassJaccardIndex"]) == pytest.approx(r2["Jaccard"])


def test_late_fusion_module_tree_equals_reference():
    """state_dict keys / shapes / default init of LateFusionModel vs the fixture written by the
    reference's own lf_model.LateFusionModel (tests/golden/make_golden.py lf)."""
    from floodplanet_code_b200.lf_model import LateFusionModel
    from floodplanet_code_b200.unet import UNetDecoder, UNetEncoder
    from floodplanet_code_b200.water_seg_model import MODELS, build_model
    fx = torch.load(GOLDEN / "lf_c4_dem1_32.pt", weights_only=False)
    cfg = fx["cfg"]
    assert MODELS["lf_model"] is LateFusionModel
    want = ["self", "in_channels", "n_classes", "lr", "log_image_iter", "to_rgb_fcn", "ignore_index",
            "optimizer_name", "feat_fusion"]
    assert list(inspect.signature(LateFusionModel.__init__).parameters) == want
    assert list(inspect.signature(UNetEncoder.__init__).parameters) == [
        "self", "n_channels", "bilinear", "base_feat_channels"]
    assert list(inspect.signature(UNetDecoder.__init__).parameters) == [
        "self", "n_classes", "bilinear", "channel_factor", "base_feat_channels"]
    torch.manual_seed(cfg["seed"])
    m = build_model("lf_model", dict(cfg["in_channels"]), cfg["n_classes"], 1e-4, 50, None,
                    cfg["ignore_index"])
    sd = m.state_dict()
    assert list(sd.keys()) == fx["state_dict_keys"]
    for k, v in fx["init_checksum"].items():
        assert float(sd[k].double().sum()) == pytest.approx(v, rel=1e-12, abs=1e-12), k
    assert sd["concat_convs.3.weight"].shape == (512, 1024, 1, 1)
    # every trainable parameter is covered by the engine's gradient slab, once
    names = m._engine.names
    assert sorted(names) == sorted(k for k, _ in m.named_parameters())
    assert len(set(names)) == len(names)
    m.load_state_dict(O.init_lf_state_dict(cfg["in_channels"], cfg["n_classes"], seed=0), strict=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.forward({"image": torch.zeros(1, 4, 32, 32), "dem": torch.zeros(1, 1, 32, 32)})
    with pytest.raises(KeyError):   # a raster without an encoder: KeyError as in the reference
        m.forward({"image": torch.zeros(1, 4, 32, 32), "slope": torch.zeros(1, 1, 32, 32)})


def test_encoder_decoder_halves_share_unet_names():
    from floodplanet_code_b200.engine import DecoderEngine, EncoderEngine, UNetEngine
    u = UNetEngine(4, 3)
    assert EncoderEngine(4).names + DecoderEngine(3).names == u.names
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        from floodplanet_code_b200.unet import UNet
        UNet(4, 3).encode(torch.zeros(1, 4, 32, 32))


def test_repack_batch_table_layout():
    """The host-built record array of `fpb200_repack_weights_batch` (include/floodplanet_b200.h): 40-byte
    records {w ptr, packed ptr, cout, cin, cin_pad, kind, first_block}, one block per output channel
    (fprop layout) or per input channel (dgrad layout), sorted by first_block."""
    import struct
    import torch
    from floodplanet_code_b200 import ops
    w1, w2 = torch.zeros(8, 4, 3, 3), torch.zeros(16, 8, 3, 3)
    o1 = torch.zeros(8, 9, 16, dtype=torch.bfloat16)
    o2 = torch.zeros(8, 9, 16, dtype=torch.bfloat16)
    table, total = ops.repack_batch_table([(w1, o1, 16, 0), (w2, o2, 8, 1)], "cpu")
    assert table.numel() == 2 * 40 and total == 8 + 8
    recs = [struct.unpack_from("<QQiiiiq", bytes(table.numpy().tobytes()), 40 * i) for i in range(2)]
    assert recs[0] == (w1.data_ptr(), o1.data_ptr(), 8, 4, 16, 0, 0)
    assert recs[1] == (w2.data_ptr(), o2.data_ptr(), 16, 8, 8, 1, 8)


def test_nccl_entry_points_validate_arguments_without_a_gpu():
    """The exchange-step entry points bind NCCL at run time (dlopen of the libnccl.so.2 torch already loaded);
    argument errors come back as status codes, nothing needs a device until a communicator is created."""
    import ctypes as C
    from floodplanet_code_b200 import capi
    lib = capi.load()
    ver = lib.fpb200_nccl_version()
    assert ver == 0 or ver >= 21800                       # 0 = no NCCL runtime in this process
    ident = (C.c_char * 128)()
    handle = C.c_void_p()
    if ver:
        assert lib.fpb200_nccl_unique_id(ident) == 0 and any(bytes(ident))
        assert lib.fpb200_nccl_comm_create(C.byref(handle), 2, 5, ident, 0) == -1      # rank outside [0, world)
        assert lib.fpb200_nccl_comm_create(C.byref(handle), 0, 0, ident, 0) == -1
    assert lib.fpb200_nccl_unique_id(None) == -6
    assert lib.fpb200_allreduce_f32(None, None, 16, 1, None) == -6                     # no communicator
    assert lib.fpb200_nccl_comm_destroy(None) == -6
    with pytest.raises(RuntimeError, match="NCCL"):
        capi.check(-6, "allreduce_f32", count=16)
