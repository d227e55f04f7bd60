# Full GPU validation: tests, smoke, default bench, reference arm (what the driver runs at round end)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "== gpu tests exit $?"; tail -n 3 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/bench_default.log | cut -c1-600
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "== ref arm exit $?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-300
