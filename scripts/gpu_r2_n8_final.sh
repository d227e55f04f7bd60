mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2y_n1.json 2> gpurun_out/r2y_n1.err; echo "n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2y_n8.json 2> gpurun_out/r2y_n8.err; echo "n8 rc=$?"
python - <<'PY'
import json
base=None
for n in (1,8):
    try:
        d=json.loads(open(f"gpurun_out/r2y_n{n}.json").read().strip().splitlines()[-1])
        base=base or d["value"]
        print(n, f"{d['value']:.1f}", d["ms_per_step"], f"eff {d['value']/(n*base):.3f}", "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"], d.get("dp_check"), (d.get("infer") or {}).get("scene_seconds"), d.get("grad_buckets_per_step"))
    except Exception as e:
        print(n, "unreadable", e); print(open(f"gpurun_out/r2y_n{n}.err").read()[-1500:])
PY
