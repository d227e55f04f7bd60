mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel|conv3x3_wgrad_kernel" -s 60 -c 8 -o gpurun_out/prof_r01_conv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_r01.log | cut -c1-400
