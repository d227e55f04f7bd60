"""The teacher-forced walker (oracle/teacher_forced.py) is itself under test, on CPU: fed the trace of a
plain-torch re-enactment of the schedule (oracle/trace_emulator.py) it must pass, and it must FAIL on
every planted fault -- a BatchNorm-backward coefficient 5 % off in ONE layer (the bug `cos > 0.75`
could not see), swapped concat halves, an upsample backward reading the wrong
half, a dropped skip-gradient add.  The GPU test (tests/test_teacher_forced_gpu.py) runs the same
walker on the CUDA path's trace."""
import pytest
import torch

from oracle import teacher_forced as TF
from oracle import trace_emulator as EM
from oracle import unet_oracle as O


def _run(fault=None, n=2, c_in=4, h=32, w=48, seed=0, frozen=False):
    from floodplanet_code_b200.engine import pad_channels, unet_conv_specs
    from floodplanet_code_b200.unet import UNet
    specs = unet_conv_specs(c_in)
    for i, s in enumerate(specs):
        s.idx = i
    sd = O.init_state_dict(c_in, 3, seed=seed)
    m = UNet(c_in, 3)
    if frozen:                       # non-trivial running statistics, as after some training
        g = torch.Generator().manual_seed(seed + 7)
        for k in sd:
            if k.endswith("running_mean"):
                sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
            elif k.endswith("running_var"):
                sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.05
            elif k.endswith("num_batches_tracked"):
                sd[k] = torch.tensor(5)
    m.load_state_dict(sd)
    if frozen:
        m.eval()
    batch = O.synthetic_batch(n, c_in, h, w, seed=seed + 1, block=8)
    logits, loss, trace = EM.emulate_traced_step(m, specs, batch, 0, pad_channels(c_in), fault=fault, frozen=frozen)
    return m, sd, batch, logits, loss, trace


def test_walker_accepts_a_correct_trace():
    m, sd, batch, logits, loss, trace = _run()
    report = TF.walk(m, sd, batch, logits, loss, trace, 0)
    assert len(report) >= 140
    assert max(e for _, e in report) < TF.GRAD_TOL
    # the emulated schedule is also a correct network: its loss equals the fp32 oracle's within 1e-2
    oloss, _, _, _ = O.training_step({k: v.clone() for k, v in sd.items()}, batch, 0, early_fusion=False)
    assert abs(float(loss) - float(oloss)) <= 1e-2 * abs(float(oloss))


@pytest.mark.parametrize("fault", ["bn_bwd_coef_layer7:0.25", "bn_bwd_coef_layer12:-0.5", "concat_halves_swapped",
                                   "upsample_bwd_reads_skip_half", "skip_gradient_dropped"])
def test_walker_rejects_planted_faults(fault):
    """At the north_star tolerances (1e-2 / 2e-2)."""
    m, sd, batch, logits, loss, trace = _run(fault)
    with pytest.raises(AssertionError):
        TF.walk(m, sd, batch, logits, loss, trace, 0)


def test_walker_at_rounding_floor_tolerance_catches_a_5_percent_coefficient_error():
    """With the tolerance tied to the bf16 rounding floor (the setting the GPU test uses, see
    TIGHT_* there) even a 5 % error in ONE BatchNorm-backward coefficient of ONE layer fails; the same
    trace without the fault passes at that tolerance."""
    m, sd, batch, logits, loss, trace = _run()
    TF.walk(m, sd, batch, logits, loss, trace, 0, fwd_tol=5e-3, grad_tol=5e-3)
    m, sd, batch, logits, loss, trace = _run("bn_bwd_coef_layer7:0.05")
    with pytest.raises(AssertionError, match="dy"):
        TF.walk(m, sd, batch, logits, loss, trace, 0, fwd_tol=5e-3, grad_tol=5e-3)


def test_walker_frozen_batchnorm_mode():
    """Eval-mode BatchNorm with autograd: the walker uses F.batch_norm(training=False), requires untouched buffers and
    compares the conv-bias gradients; it rejects a schedule that forgets the conv bias in the frozen statistics."""
    m, sd, batch, logits, loss, trace = _run(frozen=True)
    report = TF.walk(m, sd, batch, logits, loss, trace, 0, fwd_tol=5e-3, grad_tol=5e-3, frozen=True)
    assert sum(1 for k, _ in report if k.endswith("dbias")) == 18 and len(report) >= 158
    oloss, _ = O.masked_ce(O.unet_forward({k: v.clone() for k, v in sd.items()}, batch["image"], training=False),
                           batch["target"], 0)
    assert abs(float(loss) - float(oloss)) <= 1e-2 * abs(float(oloss))
    m, sd, batch, logits, loss, trace = _run("frozen_bias_dropped", frozen=True)
    with pytest.raises(AssertionError):
        TF.walk(m, sd, batch, logits, loss, trace, 0, frozen=True)
