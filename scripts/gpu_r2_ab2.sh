#!/bin/bash
# A/B on one box, alternating: launch order of a layer's dgrad / wgrad when wgrads run on the side stream
mkdir -p gpurun_out
for rep in 1 2 3; do
  for ord in ${ORDERS:-0 1}; do
    FPB200_WGRAD_AFTER_DGRAD=$ord python bench.py --steps 30 --warmup 5 --no-infer --no-cpu-baseline \
        > gpurun_out/r2ac_order${ord}_rep${rep}.json 2> gpurun_out/r2ac_order${ord}_rep${rep}.err
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2ac_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, f"{d['value']:.1f}", d.get("ms_per_step"), "e2e", d["e2e"]["value"], "clock", (d.get("clocks") or {}).get("sm_mhz"),
              (d.get("infer") or {}).get("scene_seconds"))
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-600:])
PY
