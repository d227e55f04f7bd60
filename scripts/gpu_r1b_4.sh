mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -x -q -m gpu > gpurun_out/t_k.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t_k.log
timeout 600 python scripts/conv_microbench.py --batch 64 --layers 1,0,1,15,16,17 --kinds fprop,dgrad > gpurun_out/micro_v13.log 2>&1; cat gpurun_out/micro_v13.log
timeout 600 python scripts/conv_microbench.py --batch 64 --layers 1,9,11,17 --kinds dgrad --fused-bn > gpurun_out/micro_v13f.log 2>&1; cat gpurun_out/micro_v13f.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v13.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_v13.log | cut -c1-200
timeout 900 python scripts/yardstick_cudnn.py --batch 64 > gpurun_out/yardstick.log 2>&1; echo "yardstick exit $?"; tail -4 gpurun_out/yardstick.log
