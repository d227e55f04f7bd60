mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_late_fusion_gpu.py -q -m gpu > gpurun_out/t_lf.log 2>&1; echo "lf tests exit $?"; tail -5 gpurun_out/t_lf.log
timeout 600 python scripts/bench_lf.py --batch 32 --extra dem:1 > gpurun_out/bench_lf.log 2>&1; echo "lf bench exit $?"; tail -2 gpurun_out/bench_lf.log | cut -c1-1500
timeout 300 python scripts/conv_microbench.py --batch 64 --layers 0,1 --kinds fprop,wgrad --reps 1 > gpurun_out/micro_l01.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel|conv3x3_wgrad_kernel" -c 8 -o gpurun_out/prof_l01 python scripts/conv_microbench.py --batch 64 --layers 0,1 --kinds fprop,wgrad --reps 1 > gpurun_out/ncu_l01.log 2>&1
echo "ncu exit $?"; cat gpurun_out/micro_l01.log
