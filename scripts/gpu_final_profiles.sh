# final-state evidence: conv microbench, launch list of one step, full validation
mkdir -p gpurun_out
timeout 600 python scripts/conv_microbench.py --batch 64 > gpurun_out/micro_final.log 2>&1; tail -4 gpurun_out/micro_final.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 172 -c 215 --csv --log-file gpurun_out/launches_cur.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
bash scripts/gpu_full.sh
