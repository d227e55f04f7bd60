"""Bandwidth-kernel timing at the BASELINE shapes (GPU, batch 64): ms and achieved GB/s against the
algorithmic bytes of each pass, CUDA events, best of 3."""
import sys
import torch
sys.path.insert(0, ".")
from floodplanet_code_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
def act(h, c): return torch.randn(N, h, h, c, device=dev).to(torch.bfloat16)
def f32(c): return torch.rand(c, device=dev) + 0.5
def timeit(name, fn, nbytes):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"{name:34s} {best:7.3f} ms  {nbytes / best / 1e6:8.1f} GB/s  ({nbytes/1e9:.2f} GB algorithmic)", flush=True)

for (h, c) in [(512, 64), (256, 128), (128, 256), (64, 512)]:
    y = act(h, c); a = torch.empty_like(y); da = act(h, c); dy = torch.empty_like(y)
    sc, sh, mu, istd = f32(c), f32(c), f32(c), f32(c)
    E = y.numel()
    timeit(f"bn_apply_relu {c}x{h}", lambda: ops.bn_apply_relu(y, a, sc, sh), 4 * E)
    pooled = torch.empty(N, h // 2, h // 2, c, dtype=torch.bfloat16, device=dev)
    idx = torch.empty(N, h // 2, h // 2, c, dtype=torch.uint8, device=dev)
    timeit(f"bn_apply_relu_maxpool2 {c}x{h}", lambda: ops.bn_apply_relu_maxpool2(y, a, pooled, idx, sc, sh), 4 * E + 0.75 * E)
    cat = torch.empty(N, h, h, 2 * c, dtype=torch.bfloat16, device=dev)
    timeit(f"maxpool2_bwd {c}x{h}", lambda: ops.maxpool2_bwd(pooled, idx, cat[..., :c], a), 4 * E + 0.75 * E)
    parts = torch.empty(ops.bn_bwd_rows(), 2, c, device=dev)
    timeit(f"bn_relu_bwd_reduce {c}x{h}", lambda: ops.bn_relu_bwd_reduce(da, y, sc, sh, mu, istd, parts), 4 * E)
    coef = torch.rand(2, c, device=dev)
    timeit(f"bn_relu_bwd_apply {c}x{h}", lambda: ops.bn_relu_bwd_apply(da, y, dy, sc, sh, coef), 6 * E)
    lo = act(h // 2, c)
    timeit(f"upsample_fwd {c}x{h//2}->{h}", lambda: ops.upsample2x_pad_concat_fwd(lo, cat[..., c:]), 2 * E + 0.5 * E)
    dlo = torch.empty_like(lo)
    timeit(f"upsample_bwd {c}x{h}->{h//2}", lambda: ops.upsample2x_pad_concat_bwd(cat[..., c:], dlo), 2 * E + 0.5 * E)
    del y, a, da, dy, cat, pooled, idx, lo, dlo
    torch.cuda.empty_cache()
x = act(512, 64); w = torch.randn(3, 64, device=dev); b = torch.randn(3, device=dev)
logits = torch.empty(N, 3, 512, 512, device=dev); px = N * 512 * 512
timeit("head1x1_fwd", lambda: ops.head1x1_fwd(x, w, b, logits), px * (128 + 12))
dl = torch.randn(N, 3, 512, 512, device=dev); dx = torch.empty_like(x); dw = torch.empty(3, 64, device=dev); db = torch.empty(3, device=dev)
parts = torch.empty(ops.head_bwd_rows(), 3 * 65, device=dev)
timeit("head1x1_bwd", lambda: ops.head1x1_bwd(dl, x, w, dx, dw, db, parts), px * (128 + 128 + 12))
sc64, sh64, mu64, is64 = f32(64) , f32(64) - 1.0, f32(64), f32(64)
bnp = torch.zeros(ops.head_bwd_rows(), 2, 64, device=dev)
timeit("head1x1_fwd (fused BN apply)", lambda: ops.head1x1_fwd(x, w, b, logits, sc64, sh64), px * (128 + 12))
timeit("head1x1_bwd (fused BN apply+reduce)", lambda: ops.head1x1_bwd(dl, x, w, dx, dw, db, parts, bn=(sc64, sh64, mu64, is64), bn_partials=bnp), px * (128 + 128 + 12))
tgt = (torch.rand(N, 512, 512, device=dev) < 0.42).long()
res = torch.empty(4, dtype=torch.float64, device=dev); pred = torch.empty_like(tgt); conf = torch.zeros(3, 3, dtype=torch.int64, device=dev)
cp = torch.empty(ops.ce_rows(), 4, dtype=torch.float64, device=dev)
timeit("softmax_ce_argmax_fwd", lambda: ops.softmax_ce_argmax_fwd(logits, tgt, 0, res, pred, conf, cp), px * (12 + 8 + 8))
go = torch.ones((), device=dev)
timeit("softmax_ce_bwd", lambda: ops.softmax_ce_bwd(logits, tgt, 0, res, go, dl), px * (12 + 8 + 12))
img = torch.rand(N, 4, 512, 512, device=dev)
timeit("ingest 4->16", lambda: ops.ingest([img], 16), px * (16 + 32))
# normalise + augment gather (4 fp32 channels, int64 annotation), all three transforms on every sample
from floodplanet_code_b200 import augment as G
import numpy as np
np.random.seed(0)
cfg = {"hflip": {"active": True, "likelihood": 1.0}, "vflip": {"active": True, "likelihood": 1.0},
       "rotate": {"active": True, "likelihood": 1.0, "min_rot_angle": 0, "max_rot_angle": 360}}
aug = G.DeviceAugment(cfg, "global", {"mean": np.full(4, 0.4), "std": np.full(4, 0.2)})
act = aug.sample(N)
p = G.pack_params(act, 512, 512)
th, fl, xg, yg = p["theta"].cuda(), p["flags"].cuda(), p["xgrid"].cuda(), p["ygrid"].cuda()
mean, std = aug.normalize_stats(img)
timeit("augment f32->f32 + target", lambda: ops.augment(img, tgt, th, fl, xg, yg, mean, std, want_f32=True), px * (16 + 16 + 8 + 8))
timeit("augment f32->nhwc bf16 + target", lambda: ops.augment(img, tgt, th, fl, xg, yg, mean, std, want_f32=False, c_pad=16), px * (16 + 32 + 8 + 8))
timeit("plane_mean_std (local norm)", lambda: ops.plane_mean_std(img), px * 16 * 2)
