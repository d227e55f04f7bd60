"""Per-layer conv kernel timing at the BASELINE shapes (GPU): fprop / dgrad / wgrad TFLOP/s for
each of the 18 UNet conv layers at per-GPU batch `--batch`, CUDA events, L2 flushed between
repeats by the working set itself (every layer streams >> 126 MB at batch 64).

    python scripts/conv_microbench.py [--batch 64] [--layers 1,3,16] [--kinds fprop,dgrad,wgrad]
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from floodplanet_code_b200 import ops  # noqa: E402
from floodplanet_code_b200.engine import unet_conv_specs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--layers", default="")
ap.add_argument("--kinds", default="fprop,dgrad,wgrad")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--fused-bn", action="store_true", help="dgrad with the fused BatchNorm-backward reduction")
args = ap.parse_args()

specs = unet_conv_specs(4)
layers = [int(x) for x in args.layers.split(",")] if args.layers else list(range(len(specs)))
kinds = args.kinds.split(",")
n = args.batch
tot = {k: [0.0, 0.0] for k in kinds}
for i in layers:
    s = specs[i]
    hw = args.size >> s.level
    cin = 16 if i == 0 else s.cin
    x = torch.randn(n, hw, hw, cin, device="cuda").to(torch.bfloat16)
    dy = torch.randn(n, hw, hw, s.cout, device="cuda").to(torch.bfloat16)
    w = torch.randn(s.cout, s.cin, 3, 3, device="cuda") * 0.05
    y = torch.empty(n, hw, hw, s.cout, dtype=torch.bfloat16, device="cuda")
    flops = 2.0 * n * hw * hw * s.cout * 9 * s.cin
    res = []
    for kind in kinds:
        if kind == "fprop":
            wp = ops.repack_fprop(w, cin)
            parts = torch.empty(ops.stat_rows(), 2, s.cout, device="cuda")
            fn = lambda: ops.conv3x3_fprop(x, wp, y, stat_partials=parts)
        elif kind == "dgrad":
            if i == 0:
                continue
            wd = ops.repack_dgrad(w)
            dx = torch.empty(n, hw, hw, s.cin, dtype=torch.bfloat16, device="cuda")
            if args.fused_bn and s.cin <= 512:
                yp = torch.randn(n, hw, hw, s.cin, device="cuda").to(torch.bfloat16)
                co = [torch.rand(s.cin, device="cuda") + 0.5 for _ in range(4)]
                bp = torch.empty(ops.stat_rows(), 2, s.cin, device="cuda")
                fn = lambda: ops.conv3x3_dgrad(dy, wd, dx, bn_y=yp, bn=tuple(co), bn_partials=bp)
            else:
                fn = lambda: ops.conv3x3_dgrad(dy, wd, dx)
        else:
            dw = torch.empty(s.cout, s.cin, 3, 3, device="cuda")
            ws = torch.empty(ops.wgrad_workspace_bytes(n, hw, hw, cin, s.cout) // 4, device="cuda")
            fn = lambda: ops.conv3x3_wgrad(x, dy, dw, ws, s.cin)
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res.append(f"{kind} {best:7.3f} ms {flops / best / 1e9:7.1f} TF")
        tot[kind][0] += best
        tot[kind][1] += flops
    print(f"L{i:2d} {s.cin:4d}->{s.cout:4d} @{hw:3d}  " + " | ".join(res), flush=True)
    del x, dy, y
    torch.cuda.empty_cache()
for k, (ms, fl) in tot.items():
    if ms > 0:
        print(f"total {k}: {ms:.2f} ms, {fl / ms / 1e9:.1f} TFLOP/s")
