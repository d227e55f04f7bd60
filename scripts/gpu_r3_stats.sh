#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv3x3 or full_size_conv or bn_train" > gpurun_out/r3_stats_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3_stats_tests.log
tail -n 2 gpurun_out/r3_stats_tests.log
python scripts/conv_microbench.py --batch 64 --layers 0,1,17,16,2,3,5,10 --kinds fprop 2>&1 | tee gpurun_out/r3_stats_microbench.txt
