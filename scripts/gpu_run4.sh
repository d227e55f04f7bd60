mkdir -p gpurun_out
timeout 600 python scripts/conv_microbench.py --batch 64 > gpurun_out/micro_v2.log 2>&1; echo "micro exit $?"; cat gpurun_out/micro_v2.log | tail -24
timeout 300 python scripts/conv_microbench.py --batch 16 --layers 1,3 --kinds fprop,dgrad --reps 1 > gpurun_out/plain_ncu.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -c 4 -o gpurun_out/prof_halo python scripts/conv_microbench.py --batch 16 --layers 1,3 --kinds fprop,dgrad --reps 1 > gpurun_out/ncu_halo.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_halo.log; ls -la gpurun_out/*.ncu-rep
