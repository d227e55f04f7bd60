"""Yardstick (GPU): the stock-PyTorch formulation of the same UNet training step on the same
B200 -- nn.Conv2d / BatchNorm2d / ReLU / MaxPool2d / Upsample / CrossEntropyLoss dispatching to
cuDNN / ATen -- in (a) fp32 with TF32 allowed and (b) bf16 autocast + channels_last.  This is
the "recompiled library kernels" bar of SURVEY.md section 8(d) that the hand-written path has
to beat; it is a measurement tool, not part of the product and not the parity oracle.

(c) adds torch.compile(mode="max-autotune") on top of (b) -- inductor / Triton-generated pointwise fusions
around the same cuDNN convolutions; its compilation runs in a subprocess under a time limit.

    python scripts/yardstick_cudnn.py [--batch 64] [--steps 5] [--out gpurun_out/r02_yardstick.json]
"""
import argparse
import json

import torch
import torch.nn as nn
import torch.nn.functional as F


def dc(cin, cout, mid=None):
    mid = mid or cout
    return nn.Sequential(nn.Conv2d(cin, mid, 3, padding=1), nn.BatchNorm2d(mid), nn.ReLU(inplace=True),
                         nn.Conv2d(mid, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class TorchUNet(nn.Module):
    def __init__(self, cin, ncls):
        super().__init__()
        self.inc = dc(cin, 64)
        self.downs = nn.ModuleList([dc(64, 128), dc(128, 256), dc(256, 512), dc(512, 512)])
        self.ups = nn.ModuleList([dc(1024, 256, 512), dc(512, 128, 256), dc(256, 64, 128), dc(128, 64, 64)])
        self.outc = nn.Conv2d(64, ncls, 1)

    def forward(self, x):
        feats = [self.inc(x)]
        for d in self.downs:
            feats.append(d(F.max_pool2d(feats[-1], 2)))
        u = feats[-1]
        for up, skip in zip(self.ups, reversed(feats[:-1])):
            u = F.interpolate(u, scale_factor=2, mode="bilinear", align_corners=True)
            u = up(torch.cat([skip, u], dim=1))
        return self.outc(u)


def run(mode, batch, steps, warmup):
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    model = TorchUNet(4, 3).to(dev)
    compiled = mode == "bf16_channels_last_compiled"
    if compiled:
        mode = "bf16_channels_last"
    if mode == "bf16_channels_last":
        model = model.to(memory_format=torch.channels_last)
    fwd = torch.compile(model, mode="max-autotune") if compiled else model
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    lossf = nn.CrossEntropyLoss(ignore_index=0)
    x = torch.rand(batch, 4, 512, 512, device=dev)
    if mode == "bf16_channels_last":
        x = x.contiguous(memory_format=torch.channels_last)
    t = (torch.rand(batch, 16, 16, device=dev) > 0.58).long().repeat_interleave(32, 1).repeat_interleave(32, 2)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_channels_last")):
            out = fwd(x)
        loss = lossf(out.float(), t)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"mode": mode + ("+torch.compile(max-autotune)" if compiled else ""), "batch": batch, "ms_per_step": ms, "chips_per_s": batch / ms * 1e3,
            "tflops": 959.42e9 * batch / ms / 1e9, "loss": float(loss),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default=None, help="also write all rows as one JSON document")
    ap.add_argument("--only", default=None, help="(internal) run a single mode in this process")
    ap.add_argument("--compile-timeout", type=int, default=900)
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    rows = []
    modes = (args.only,) if args.only else ("bf16_channels_last", "fp32_tf32")
    for mode in modes:
        b = args.batch
        while b >= 8:
            try:
                torch.cuda.reset_peak_memory_stats()
                rows.append(run(mode, b, args.steps, args.warmup))
                print(json.dumps(rows[-1]), flush=True)
                break
            except torch.OutOfMemoryError:
                torch.cuda.empty_cache()
                b //= 2
    if not args.only:
        # torch.compile row in a subprocess: compilation + autotuning of 18 conv layers x fwd/bwd can take many
        # minutes and must not take the whole yardstick down with it
        import subprocess
        import sys
        cmd = [sys.executable, __file__, "--batch", str(args.batch), "--steps", str(args.steps), "--warmup",
               str(args.warmup), "--only", "bf16_channels_last_compiled"]
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=args.compile_timeout)
            got = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
            rows += got if got else [{"mode": "bf16_channels_last+torch.compile(max-autotune)",
                                      "error": (res.stderr or "no output")[-400:]}]
        except subprocess.TimeoutExpired:
            rows.append({"mode": "bf16_channels_last+torch.compile(max-autotune)",
                         "error": f"compilation did not finish within {args.compile_timeout} s"})
        print(json.dumps(rows[-1]), flush=True)
        if args.out:
            doc = {"what": "stock PyTorch formulation of the same UNet train step (nn.Conv2d / BatchNorm2d / ... -> cuDNN / "
                           "ATen, fused torch Adam), same B200, same synthetic batch; the library bar the hand-written "
                           "path is measured against (SURVEY.md 8d)",
                   "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
                   "gpu": torch.cuda.get_device_name(0), "rows": rows}
            open(args.out, "w").write(json.dumps(doc, indent=1))
