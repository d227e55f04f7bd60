mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 205 -c 215 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
