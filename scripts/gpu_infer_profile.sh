mkdir -p gpurun_out
timeout 600 python scripts/bench_infer.py --reps 1 > gpurun_out/infer_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 480 -c 480 --csv --log-file gpurun_out/launches_infer.csv python scripts/bench_infer.py --reps 1 > gpurun_out/ncu_infer.log 2>&1
echo "ncu exit $?"
