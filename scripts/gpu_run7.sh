mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "wgrad" --tb=short -p no:cacheprovider > gpurun_out/t_wgrad7.log 2>&1; echo "== wgrad tests exit $?"; tail -n 12 gpurun_out/t_wgrad7.log
timeout 600 python scripts/conv_microbench.py --batch 64 --kinds wgrad > gpurun_out/micro_wg.log 2>&1; echo "micro exit $?"; cat gpurun_out/micro_wg.log | tail -22
