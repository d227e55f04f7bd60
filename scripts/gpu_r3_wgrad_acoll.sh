#!/bin/bash
# A/B of the A collector in the MODE_X_SHIFT NB = 2 wgrad (FPB200_WGRAD_ACOLL), alternating on one box
mkdir -p gpurun_out
FPB200_WGRAD_ACOLL=1 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "wgrad or full_size_conv" > gpurun_out/r3_acoll_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3_acoll_tests.log
tail -n 2 gpurun_out/r3_acoll_tests.log
for m in 0 1 0 1; do
  echo "== FPB200_WGRAD_ACOLL=$m" | tee -a gpurun_out/r3_acoll_microbench.txt
  FPB200_WGRAD_ACOLL=$m python scripts/conv_microbench.py --batch 64 --layers 3,5,7,10,12,14 --kinds wgrad 2>&1 | tee -a gpurun_out/r3_acoll_microbench.txt
done
