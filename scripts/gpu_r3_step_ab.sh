#!/bin/bash
# in-step A/B of the weight-stationary halo kernel (alternating, same box), then the full GPU suite
mkdir -p gpurun_out
for rep in 1 2; do
  for m in 1 3; do
    FPB200_HALO_WS=$m python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-infer > gpurun_out/r3_step_ws${m}_rep${rep}.json 2> gpurun_out/r3_step_ws${m}_rep${rep}.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/r3_step_ws${m}_rep${rep}.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("ws=${m} rep=${rep}", round(d["value"], 1), "chips/s", round(d["ms_per_step"], 2), "ms  e2e", round(d["e2e"]["value"], 1), "clk", d["clocks"]["sm_mhz"],
      {k: (round(v["ms_per_step"], 2), round(v["achieved"])) for k, v in r["families"].items()}, r["worst_layers"][:3])
PY
  done
done


