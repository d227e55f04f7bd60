#!/bin/bash
# round 2, 1 GPU: changed tests, headline bench with the inference block, stock-Adam and default-workload variants
mkdir -p gpurun_out
python -m pytest tests/test_inference.py tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r2b_bench_n1.err
python bench.py --steps 10 --warmup 3 --optimizer torch --no-infer --no-cpu-baseline > gpurun_out/r2b_bench_n1_torchadam.json 2> gpurun_out/r2b_bench_n1_torchadam.err
for opt in fused torch; do
  python bench.py --steps 30 --warmup 5 --batch 10 --size 300 --optimizer $opt --no-infer --no-cpu-baseline > gpurun_out/r2b_bench_b10_s300_${opt}.json 2> gpurun_out/r2b_bench_b10_s300_${opt}.err
done
python bench.py --steps 30 --warmup 5 --batch 10 --size 300 --graph --no-infer --no-cpu-baseline > gpurun_out/r2b_bench_b10_s300_graph.json 2> gpurun_out/r2b_bench_b10_s300_graph.err
for f in gpurun_out/r2b_*.err gpurun_out/r2b_tests.log; do echo "== $f"; tail -n 4 $f; done
for f in gpurun_out/r2b_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","cuda_graph","optimizer_impl")}, d["e2e"], d.get("clocks"))
    if d.get("infer"): print(d["infer"])
except Exception as e: print("unreadable", e)
PY
done
