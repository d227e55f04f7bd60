#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "head" > gpurun_out/r3_head_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3_head_tests.log
tail -n 3 gpurun_out/r3_head_tests.log
python scripts/elem_microbench.py 64 2>&1 | grep -i "head\|upsample\|maxpool2_bwd 64\|bn_relu_bwd" | tee gpurun_out/r3_head_microbench.txt
