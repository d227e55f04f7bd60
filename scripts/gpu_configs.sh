# BASELINE configs[3] (early fusion, wide-input first conv) and configs[4] (10240^2 scene inference) on the current build
mkdir -p gpurun_out
for c in 16 6 21; do timeout 600 python bench.py --channels $c --no-cpu-baseline --steps 5 > gpurun_out/bench_ef_c$c.log 2>&1; echo "ef c=$c exit $?"; tail -n 1 gpurun_out/bench_ef_c$c.log | cut -c1-160; done
timeout 600 python scripts/bench_infer.py > gpurun_out/infer_n1.log 2>&1; echo "infer exit $?"; tail -n 1 gpurun_out/infer_n1.log | cut -c1-400
timeout 600 python scripts/bench_infer.py --stride 256 > gpurun_out/infer_n1_s256.log 2>&1; echo "infer s256 exit $?"; tail -n 1 gpurun_out/infer_n1_s256.log | cut -c1-400
timeout 600 python scripts/bench_lf.py > gpurun_out/bench_lf.log 2>&1; echo "lf exit $?"; tail -n 2 gpurun_out/bench_lf.log | cut -c1-300
