mkdir -p gpurun_out
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 190 -c 230 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
timeout 600 python scripts/conv_microbench.py --batch 64 > gpurun_out/micro_final.log 2>&1; tail -3 gpurun_out/micro_final.log
timeout 900 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_final.log | cut -c1-200
