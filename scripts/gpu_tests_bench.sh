# GPU tests + one bench line (no CPU baseline, short)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "== gpu tests exit $?"; tail -n 5 gpurun_out/t_all.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_cur.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/bench_cur.log | cut -c1-200
timeout 900 python bench.py --no-cpu-baseline --graph > gpurun_out/bench_graph.log 2>&1; echo "== bench graph exit $?"; tail -n 1 gpurun_out/bench_graph.log | cut -c1-200
