mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_late_fusion_gpu.py tests/test_kernels_gpu.py -x -q -m gpu -k "conv1x1 or layout or late_fusion or fusion_conv or encode or encoder" > gpurun_out/t_new.log 2>&1; echo "new tests exit $?"; tail -25 gpurun_out/t_new.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all tests exit $?"; tail -8 gpurun_out/t_all.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1b.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_r1b.log | cut -c1-400
