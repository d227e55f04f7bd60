#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
cat > /tmp/uponly.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from floodplanet_code_b200 import ops
N = 64; dev = "cuda"; h = 512; c = 64
cat = torch.randn(N, h, h, 2 * c, device=dev).to(torch.bfloat16)
dlo = torch.empty(N, h // 2, h // 2, c, dtype=torch.bfloat16, device=dev)
for _ in range(2):
    ops.upsample2x_pad_concat_bwd(cat[..., c:], dlo)
torch.cuda.synchronize()
PY
python /tmp/uponly.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:upsample2x_pad_bwd_kernel -s 1 -c 1 -o /tmp/ncu/up -f python /tmp/uponly.py > gpurun_out/r3_up_ncu.log 2>&1
ncu -i /tmp/ncu/up.ncu-rep --page raw --csv > gpurun_out/r3_up_raw.csv 2>/dev/null
ncu -i /tmp/ncu/up.ncu-rep --page source --csv > gpurun_out/r3_up_source.csv 2>/dev/null
ls -la gpurun_out/r3_up_*
