#!/bin/bash
# round 2 evidence (1 GPU): launch list of one training step, ncu --set full of the CURRENT conv kernel modes,
# cuDNN / torch.compile yardstick.  Every ncu command runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
# tcgen05 operand-path microbenchmark (built on the box: gpurun_out/ does not travel)
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I floodplanet_code_b200/csrc scripts/umma_microbench.cu \
    -o gpurun_out/umma_microbench > gpurun_out/r2p_umma_build.log 2>&1 && timeout 120 gpurun_out/umma_microbench > gpurun_out/r2p_umma_microbench.txt 2>&1
echo "umma microbench rc=$?"; cat gpurun_out/r2p_umma_microbench.txt
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-infer"
timeout 600 $B > gpurun_out/r2p_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -s 225 -c 270 --csv --log-file gpurun_out/r2p_launches.csv $B > gpurun_out/r2p_ncu_list.log 2>&1
echo "launch list rc=$?"
M="python scripts/conv_microbench.py --batch 64 --layers 0,1,2,3,10,16 --kinds fprop,wgrad --reps 1"
timeout 600 $M > gpurun_out/r2p_micro_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel|conv3x3_wgrad_kernel" \
    -c 24 -o gpurun_out/r2p_conv_modes $M > gpurun_out/r2p_ncu_modes.log 2>&1
echo "ncu modes rc=$?"
ncu -i gpurun_out/r2p_conv_modes.ncu-rep --page raw --csv > gpurun_out/r2p_conv_modes_raw.csv 2>/dev/null
rm -f gpurun_out/r2p_conv_modes.ncu-rep      # 77 MB: gpurun merges back at most 64 MiB; the raw CSV carries every metric
python scripts/conv_microbench.py --batch 64 > gpurun_out/r2p_conv_microbench_b64.txt 2>&1
[ "$SKIP_YARDSTICK" = "1" ] || timeout 1500 python scripts/yardstick_cudnn.py --batch 64 --steps 5 --compile-timeout 700 --out gpurun_out/r2p_yardstick.json \
    > gpurun_out/r2p_yardstick.log 2>&1
echo "yardstick rc=$?"; tail -n 4 gpurun_out/r2p_yardstick.log | cut -c1-300
tail -n 3 gpurun_out/r2p_conv_microbench_b64.txt
ls -la gpurun_out/r2p_*
