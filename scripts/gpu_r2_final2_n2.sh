#!/bin/bash
# round 2, last session: 2-GPU check of the final code -- hardware gradient-parity tests, then the driver's launch at N = 2
mkdir -p gpurun_out
python -m pytest tests/test_parallel_gpu.py -q -m gpu > gpurun_out/r2h_n2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2h_n2_tests.log
tail -n 3 gpurun_out/r2h_n2_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2h_bench_n2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "gpu_launches", "grad_buckets_per_step", "dp_check")}, d.get("e2e"), d.get("clocks"))
if d.get("infer"): print("  infer", {k: d["infer"].get(k) for k in ("scene_seconds", "tiles_per_sec", "tflops")})
PY
