mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for k in "fprop_and_stats" "affine_relu" "dgrad" "wgrad" "not conv3x3 and not wgrad"; do
  name=$(echo "$k" | tr ' ' '_')
  timeout 400 python -m pytest tests/test_kernels_gpu.py -q -k "$k" --tb=short -p no:cacheprovider > gpurun_out/t_$name.log 2>&1
  echo "== $k : exit $?"
  tail -n 25 gpurun_out/t_$name.log
done
