mkdir -p gpurun_out
for k in "halo-" "halo_bo1" "per_tap"; do
  timeout 400 python -m pytest tests/test_kernels_gpu.py -q -k "conv3x3 and $k" --tb=line -p no:cacheprovider > gpurun_out/t_conv_$k.log 2>&1
  echo "== conv tests [$k] exit $?"; tail -n 12 gpurun_out/t_conv_$k.log
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2.log 2>&1; echo "== bench exit $?"; tail -n 2 gpurun_out/bench_v2.log
