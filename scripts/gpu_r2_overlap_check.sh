python -m pytest tests -q -m gpu -x > gpurun_out/r2h_tests.log 2>&1; echo rc=$? >> gpurun_out/r2h_tests.log; tail -n 6 gpurun_out/r2h_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_n2.json 2> gpurun_out/r2h_n2.err; echo "n2 rc=$?"
FPB200_OVERLAP_WGRAD=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-infer > gpurun_out/r2h_n2_serial.json 2> gpurun_out/r2h_n2_serial.err; echo "n2 serial rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2h_n2.json", "gpurun_out/r2h_n2_serial.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, f"{d['value']:.1f}", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"], d["dp_check"], (d.get("infer") or {}).get("scene_seconds"))
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
