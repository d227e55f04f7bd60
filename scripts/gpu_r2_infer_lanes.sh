mkdir -p gpurun_out
python -m pytest tests/test_inference.py -q -m gpu 2>&1 | tail -n 2
for rep in 1 2; do
  for lanes in 1 2; do
    FPB200_INFER_LANES=$lanes python bench.py --infer-only > gpurun_out/r2l_lanes${lanes}_rep${rep}.json 2> gpurun_out/r2l_lanes${lanes}_rep${rep}.err
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2l_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); i = d["infer"]
        print(f, i["scene_seconds"], i["scene_seconds_all_passes"], i["tflops"], d["clocks"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-800:])
PY
