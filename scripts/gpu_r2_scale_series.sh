#!/bin/bash
# the driver's scaling protocol, run by the builder on ONE 8-GPU box: N = 1, 2, 4, 8 back to back, default flags
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2x_scale_n$N.json 2> gpurun_out/r2x_scale_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
       bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2x_scale_n$N.json 2> gpurun_out/r2x_scale_n$N.err
  fi
  echo "N=$N rc=$?" >> gpurun_out/r2x_scale_n$N.err
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/r2x_scale_n{n}.json").read().strip().splitlines()[-1])
        base = base or d["value"]
        inf = d.get("infer") or {}
        print(n, f"{d['value']:.1f} chips/s {d['ms_per_step']:.2f} ms  eff {d['value'] / (n * base):.3f}  e2e {d['e2e']['value']:.1f}  "
                 f"clock {d['clocks']['sm_mhz']}  infer {inf.get('scene_seconds')} s {inf.get('tiles_per_sec')} tiles/s  dp_check {d.get('dp_check')}")
    except Exception as e:
        print(n, "unreadable", e); print(open(f"gpurun_out/r2x_scale_n{n}.err").read()[-800:])
PY
