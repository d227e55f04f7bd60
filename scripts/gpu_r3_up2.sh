#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/up2.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from floodplanet_code_b200 import ops
N = 64; dev = "cuda"
def timeit(name, fn, nbytes):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"{name:44s} {best:7.3f} ms  {nbytes / best / 1e6:8.1f} GB/s", flush=True)
for (h, c) in [(512, 64), (256, 128)]:
    E = N * h * h * c
    cat = torch.randn(N, h, h, 2 * c, device=dev).to(torch.bfloat16)
    dense = torch.randn(N, h, h, c, device=dev).to(torch.bfloat16)
    dlo = torch.empty(N, h // 2, h // 2, c, dtype=torch.bfloat16, device=dev)
    timeit(f"upsample_bwd {c}ch {h}->{h//2} upper half of concat", lambda: ops.upsample2x_pad_concat_bwd(cat[..., c:], dlo), 2.5 * E)
    timeit(f"upsample_bwd {c}ch {h}->{h//2} lower half of concat", lambda: ops.upsample2x_pad_concat_bwd(cat[..., :c], dlo), 2.5 * E)
    timeit(f"upsample_bwd {c}ch {h}->{h//2} dense input", lambda: ops.upsample2x_pad_concat_bwd(dense, dlo), 2.5 * E)
    lo = torch.randn(N, h // 2, h // 2, c, device=dev).to(torch.bfloat16)
    timeit(f"upsample_fwd {c}ch into upper half of concat", lambda: ops.upsample2x_pad_concat_fwd(lo, cat[..., c:]), 2.5 * E)
    timeit(f"upsample_fwd {c}ch into dense output", lambda: ops.upsample2x_pad_concat_fwd(lo, dense), 2.5 * E)
    y = torch.randn(N, h, h, c, device=dev).to(torch.bfloat16); a = torch.empty_like(y)
    sc = torch.rand(c, device=dev) + 0.5
    timeit(f"bn_apply_relu {c}ch dense -> dense", lambda: ops.bn_apply_relu(y, a, sc, sc), 4 * E)
    timeit(f"bn_apply_relu {c}ch dense -> half of concat", lambda: ops.bn_apply_relu(y, cat[..., :c], sc, sc), 4 * E)
PY
python /tmp/up2.py 2>&1 | tee gpurun_out/r3_up2.txt
