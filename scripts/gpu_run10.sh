mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_inference.py tests/test_kernels_gpu.py tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_10.log 2>&1; echo "== tests exit $?"; tail -n 12 gpurun_out/t_10.log
timeout 600 python scripts/bench_infer.py > gpurun_out/infer_n1.log 2>&1; echo "== infer exit $?"; tail -n 2 gpurun_out/infer_n1.log
timeout 600 python scripts/bench_infer.py --stride 256 --reps 1 > gpurun_out/infer_n1_s256.log 2>&1; echo "== infer s256 exit $?"; tail -n 2 gpurun_out/infer_n1_s256.log
