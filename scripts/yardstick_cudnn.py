"""Yardstick (GPU): the stock-PyTorch formulation of the same UNet training step on the same
B200 -- nn.Conv2d / BatchNorm2d / ReLU / MaxPool2d / Upsample / CrossEntropyLoss dispatching to
cuDNN / ATen -- in (a) fp32 with TF32 allowed and (b) bf16 autocast + channels_last.  This is
the "recompiled library kernels" bar of SURVEY.md section 8(d) that the hand-written path has
to beat; it is a measurement tool, not part of the product and not the parity oracle.

    python scripts/yardstick_cudnn.py [--batch 64] [--steps 5]
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
    if mode == "bf16_channels_last":
        model = model.to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    lossf = nn.CrossEntropyLoss(ignore_index=0)
    x = torch.rand(batch, 4, 512, 512, device=dev)
    if mode == "bf16_channels_last":
        x = x.contiguous(memory_format=torch.channels_last)
    t = (torch.rand(batch, 16, 16, device=dev) > 0.58).long().repeat_interleave(32, 1).repeat_interleave(32, 2)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_channels_last")):
            out = model(x)
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
    return {"mode": mode, "batch": batch, "ms_per_step": ms, "chips_per_s": batch / ms * 1e3,
            "tflops": 959.42e9 * batch / ms / 1e9, "loss": float(loss),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    for mode in ("bf16_channels_last", "fp32_tf32"):
        b = args.batch
        while b >= 8:
            try:
                torch.cuda.reset_peak_memory_stats()
                print(json.dumps(run(mode, b, args.steps, args.warmup)), flush=True)
                break
            except torch.OutOfMemoryError:
                torch.cuda.empty_cache()
                b //= 2
