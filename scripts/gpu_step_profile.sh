# ncu launch list of ONE training step (after the same command ran clean without ncu).  The window is a
# little wider than a step; scripts/ncu_launch_table.py --step trims it to [ingest_kernel .. repack_batch_kernel]
mkdir -p gpurun_out
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 225 -c 260 --csv --log-file gpurun_out/launches_cur.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
