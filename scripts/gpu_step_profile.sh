# launch list of one training step (after the same command ran clean without ncu) + elementwise microbench
mkdir -p gpurun_out
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 190 -c 230 --csv --log-file gpurun_out/launches_cur.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
timeout 300 python scripts/elem_microbench.py 64 > gpurun_out/elem_cur.log 2>&1; tail -4 gpurun_out/elem_cur.log
