#!/bin/bash
# round 2, last session: the driver's launch at N = 8 with the final code
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n8.json 2> gpurun_out/r2i_bench_n8.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2i_bench_n8.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "gpu_launches", "grad_buckets_per_step", "dp_check")}, d.get("e2e"), d.get("clocks"))
if d.get("infer"): print("  infer", {k: d["infer"].get(k) for k in ("scene_seconds", "tiles_per_sec", "tflops")})
PY
