#!/bin/bash
# round 2, 8 GPUs, second pass: per-step jitter attribution (replica probe) and the default DP configuration at 30 steps
N=${1:-8}
mkdir -p gpurun_out
run() { name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2n_n${N}_${name}.json 2> gpurun_out/r2n_n${N}_${name}.err
  echo "$name rc=$?" >> gpurun_out/r2n_n${N}_${name}.err; }
run replicas --no-allreduce --no-infer
run capi --no-infer
run capi_onebucket --bucket-mb 128 --no-infer
python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-infer > gpurun_out/r2n_n${N}_samebox_n1.json 2> gpurun_out/r2n_n${N}_samebox_n1.err
for f in gpurun_out/r2n_n${N}_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","grad_buckets_per_step","exchange_step")}, "clocks", d.get("clocks"))
    print(" jitter", json.dumps(d.get("step_jitter"))[:1200])
    if d.get("allreduce"): print(" allreduce", json.dumps(d["allreduce"])[:900])
except Exception as e: print("unreadable", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
