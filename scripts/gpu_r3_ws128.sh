#!/bin/bash
# A/B of weight-stationary MMAs in the 128-output-channel halo kernels (FPB200_HALO_WS bit 1), alternating on one box
mkdir -p gpurun_out
FPB200_HALO_WS=3 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv3x3 or full_size_conv" > gpurun_out/r3_ws128_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3_ws128_tests.log
tail -n 3 gpurun_out/r3_ws128_tests.log
for m in 1 3 1 3; do
  echo "== FPB200_HALO_WS=$m" | tee -a gpurun_out/r3_ws128_microbench.txt
  FPB200_HALO_WS=$m python scripts/conv_microbench.py --batch 64 --layers 2,3,5,7,10,12,14 --kinds fprop,dgrad 2>&1 | tee -a gpurun_out/r3_ws128_microbench.txt
done
