#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
cat > /tmp/headonly.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from floodplanet_code_b200 import ops
N = 64; dev = "cuda"
x = torch.randn(N, 512, 512, 64, device=dev).to(torch.bfloat16); w = torch.randn(3, 64, device=dev)
dl = torch.randn(N, 3, 512, 512, device=dev); dx = torch.empty_like(x); dw = torch.empty(3, 64, device=dev); db = torch.empty(3, device=dev)
parts = torch.empty(ops.head_bwd_rows(), 3 * 65, device=dev)
f32 = lambda c: torch.rand(c, device=dev) + 0.5
sc, sh, mu, istd = f32(64), f32(64) - 1.0, f32(64), f32(64)
bnp = torch.zeros(ops.head_bwd_rows(), 2, 64, device=dev)
for _ in range(2):
    ops.head1x1_bwd(dl, x, w, dx, dw, db, parts, bn=(sc, sh, mu, istd), bn_partials=bnp)
torch.cuda.synchronize()
PY
python /tmp/headonly.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:head1x1_bwd_kernel -s 1 -c 1 -o /tmp/ncu/head -f python /tmp/headonly.py > gpurun_out/r3_head_ncu.log 2>&1
ncu -i /tmp/ncu/head.ncu-rep --page raw --csv > gpurun_out/r3_head_raw.csv 2>/dev/null
ncu -i /tmp/ncu/head.ncu-rep --page source --csv > gpurun_out/r3_head_source.csv 2>/dev/null
ls -la gpurun_out/r3_head_*
