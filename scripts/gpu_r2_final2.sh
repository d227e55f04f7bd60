#!/bin/bash
# round 2, last session: final check on one GPU the way the driver runs things (smoke, full GPU suite, reference arm,
# headline bench), then the ncu launch list of one training step (after the same command exited 0 without ncu)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2k_smoke.log
python -m pytest tests -q -m gpu > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2k_tests.log

( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err
for f in gpurun_out/r2k_smoke.log gpurun_out/r2k_tests.log gpurun_out/r2k_bench_n1.err; do echo "== $f"; tail -n 6 $f; done
python - <<'PY'
import json
for f in ("gpurun_out/r2k_bench_n1.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "reference_sample")}, d.get("e2e"), d.get("clocks"))
        if d.get("roofline"): print("  roofline", {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "traffic", "share_of_step")})
        if d.get("cpu_baseline"): print("  cpu", d["cpu_baseline"])
        if d.get("infer"): print("  infer", {k: d["infer"][k] for k in ("scene_seconds", "tiles_per_sec", "tflops", "frac_of_sustained_bf16_peak_per_gpu", "h2d_bytes_per_scene", "d2h_bytes_per_scene")})
    except Exception as e:
        print(f, "unreadable", e)
PY
