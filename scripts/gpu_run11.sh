mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "conv3x3 or wgrad" --tb=line -p no:cacheprovider > gpurun_out/t_conv11.log 2>&1; echo "== conv tests exit $?"; tail -n 3 gpurun_out/t_conv11.log
timeout 600 python scripts/conv_microbench.py --batch 64 > gpurun_out/micro_v5.log 2>&1; echo "micro exit $?"; cat gpurun_out/micro_v5.log | tail -22
