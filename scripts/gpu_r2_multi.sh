#!/bin/bash
# round 2, N GPUs (N = $1): data-parallel gradient parity test + bench variants that attribute the scaling loss
N=${1:-2}
mkdir -p gpurun_out
run() { # name, extra args
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2m_n${N}_${name}.json 2> gpurun_out/r2m_n${N}_${name}.err
  echo "$name rc=$?" >> gpurun_out/r2m_n${N}_${name}.err
}
if [ "$N" = "2" ]; then
  python -m pytest tests/test_parallel_gpu.py -q -m gpu -s > gpurun_out/r2m_parallel_test.log 2>&1
  echo "rc=$?" >> gpurun_out/r2m_parallel_test.log
  tail -n 5 gpurun_out/r2m_parallel_test.log
fi
run capi
if [ "$N" != "2" ]; then
  python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-infer > gpurun_out/r2m_n${N}_samebox_n1.json 2> gpurun_out/r2m_n${N}_samebox_n1.err
  run torch --transport torch --no-infer
  run capi_cta8 --nccl-max-ctas 8 --no-infer
  run capi_onebucket --bucket-mb 128 --no-infer
  run replicas --no-allreduce --no-infer
fi
for f in gpurun_out/r2m_n${N}_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","per_rank_ms_per_step","grad_buckets_per_step","dp_check","exchange_step")})
    print(" e2e", d["e2e"]["value"], "clocks", d.get("clocks"))
    if d.get("allreduce"): print(" allreduce", json.dumps(d["allreduce"])[:1500])
    if d.get("infer"): print(" infer", {k:d["infer"][k] for k in ("scene_seconds","tiles_per_sec","tflops","frac_of_sustained_bf16_peak_per_gpu")})
except Exception as e: print("unreadable", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
