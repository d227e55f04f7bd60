#!/bin/bash
# round 3 (second half of round 2): weight-stationary MMAs in the 64-output-channel halo kernels -- parity, then A/B
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv3x3 or full_size_conv or conv1x1" > gpurun_out/r3_ws_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3_ws_tests.log
tail -n 4 gpurun_out/r3_ws_tests.log
for m in 0 1 0 1; do
  echo "== FPB200_HALO_WS=$m" | tee -a gpurun_out/r3_ws_microbench.txt
  FPB200_HALO_WS=$m python scripts/conv_microbench.py --batch 64 --layers 0,1,2,15,16,17,3,10 --kinds fprop,dgrad 2>&1 | tee -a gpurun_out/r3_ws_microbench.txt
done
