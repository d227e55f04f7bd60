#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "bn_relu_backward or bn_train or full_size_pool" > gpurun_out/r3_reduce_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3_reduce_tests.log
tail -n 2 gpurun_out/r3_reduce_tests.log
python scripts/elem_microbench.py 64 2>&1 | grep -i "bn_relu_bwd_reduce" | tee gpurun_out/r3_reduce_microbench.txt
