#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_teacher_forced_gpu.py -q -m gpu -k "upsample or bilinear or teacher" > gpurun_out/r3_up_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3_up_tests.log
tail -n 5 gpurun_out/r3_up_tests.log
python scripts/elem_microbench.py 64 2>&1 | grep -i "upsample" | tee gpurun_out/r3_up_microbench.txt
