mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider -k "not conv3x3 and not wgrad" > gpurun_out/t_model.log 2>&1
echo "== model tests exit $?"; tail -n 40 gpurun_out/t_model.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 5 gpurun_out/smoke.log
timeout 600 python bench.py --batch 8 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_b8.log 2>&1; echo "== bench b8 exit $?"; tail -n 3 gpurun_out/bench_b8.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b64.log 2>&1; echo "== bench b64 exit $?"; tail -n 3 gpurun_out/bench_b64.log
