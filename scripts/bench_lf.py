"""Late-fusion training-step timing (GPU): LateFusionModel with `--modalities` encoders on
synthetic 512x512 chips, fwd + masked CE + bwd + fused Adam, CUDA events.  Also times the
pointwise fusion kernels in isolation.  Not the headline bench (bench.py); SURVEY.md 8(f) rank 4.

    python scripts/bench_lf.py [--batch 32] [--extra dem:1,slope:1] [--steps 5]
"""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from floodplanet_code_b200 import ops  # noqa: E402
from floodplanet_code_b200.engine import FEAT_CH, decoder_conv_specs, encoder_conv_specs  # noqa: E402
from floodplanet_code_b200.optim import FusedAdam  # noqa: E402
from floodplanet_code_b200.water_seg_model import build_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--extra", default="dem:1")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()

in_ch = {"ms_image": 4}
for item in filter(None, args.extra.split(",")):
    k, c = item.split(":")
    in_ch[k] = int(c)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_model("lf_model", in_ch, 3, 1e-4, 50, None, 0).to(dev)
opt = FusedAdam(model, 1e-4)
n, s = args.batch, args.size
g = torch.Generator(device=dev).manual_seed(0)
batch = {"image": torch.rand(n, 4, s, s, generator=g, device=dev)}
for k, c in in_ch.items():
    if k != "ms_image":
        batch[k] = torch.rand(n, c, s, s, generator=g, device=dev)
batch["target"] = (torch.rand(n, s // 32, s // 32, generator=g, device=dev) > 0.58).long() \
    .repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()


def step():
    opt.zero_grad()
    loss = model.training_step(batch, 0)
    loss.backward()
    opt.step()
    return loss


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps

k = len(in_ch)
px = [(s >> l) ** 2 for l in range(5)]
conv_fwd = sum(2.0 * px[sp.level] * sp.cout * 9 * sp.cin for c in in_ch.values() for sp in encoder_conv_specs(c))
conv_fwd += sum(2.0 * px[sp.level] * sp.cout * 9 * sp.cin for sp in decoder_conv_specs())
pw_fwd = sum(2.0 * px[l] * FEAT_CH[l] * FEAT_CH[l] * k for l in range(5))
flops_step = 3.0 * (conv_fwd + pw_fwd) * n          # fwd + dgrad + wgrad (first-layer dgrad not subtracted)
out = {"workload": f"late fusion {in_ch}, batch {n}, {s}x{s}", "ms_per_step": ms, "chips_per_s": n / ms * 1e3,
       "approx_tflops": flops_step / ms / 1e9, "loss": float(loss), "gflop_per_chip_fwd": (conv_fwd + pw_fwd) / 1e9,
       "pointwise_share_of_flops": pw_fwd / (conv_fwd + pw_fwd)}

# the fusion kernels in isolation
pw = {}
for l in range(5):
    hw = s >> l
    fs = FEAT_CH[l]
    x = torch.randn(n, hw, hw, fs * k, device=dev).to(torch.bfloat16)
    dy = torch.randn(n, hw, hw, fs, device=dev).to(torch.bfloat16)
    w = torch.randn(fs, fs * k, 1, 1, device=dev) * 0.05
    y = torch.empty(n, hw, hw, fs, dtype=torch.bfloat16, device=dev)
    dx = torch.empty_like(x)
    dw = torch.empty_like(w)
    ws = torch.empty(ops.conv1x1_wgrad_workspace_bytes(n, hw, hw, fs * k, fs) // 4, device=dev)
    wp, wt = ops.repack_1x1(w, False), ops.repack_1x1(w, True)
    one, b = torch.ones(fs, device=dev), torch.zeros(fs, device=dev)
    fl = 2.0 * n * hw * hw * fs * fs * k
    byt = n * hw * hw * (fs * k + fs) * 2.0
    for name, fn in (("fprop", lambda: ops.conv1x1(x, wp, y, one, b)), ("dgrad", lambda: ops.conv1x1(dy, wt, dx)),
                     ("wgrad", lambda: ops.conv1x1_wgrad(x, dy, dw, ws))):
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); fn(); a1.record(); torch.cuda.synchronize()
            best = min(best, a0.elapsed_time(a1))
        pw[f"L{l}:{name}"] = {"ms": round(best, 3), "TF": round(fl / best / 1e9, 1), "GB/s": round(byt / best / 1e6, 1)}
    del x, dy, y, dx
out["pointwise"] = pw
print(json.dumps(out))
