#!/bin/bash
# A/B of the tap-boundary form of the halo issue loop (FPB200_HALO_ISSUE 0 / 1 / 2), alternating on one box
mkdir -p gpurun_out
for m in 1 2; do
FPB200_HALO_ISSUE=$m python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv3x3 or full_size_conv" > gpurun_out/r3_issue${m}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3_issue${m}_tests.log
tail -n 2 gpurun_out/r3_issue${m}_tests.log
done
for m in 0 1 2 0 1 2; do
  echo "== FPB200_HALO_ISSUE=$m" | tee -a gpurun_out/r3_issue_microbench.txt
  FPB200_HALO_ISSUE=$m python scripts/conv_microbench.py --batch 64 --layers 1,16,17,3,5,10,12 --kinds fprop,dgrad 2>&1 | tee -a gpurun_out/r3_issue_microbench.txt
done
