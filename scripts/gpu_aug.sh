mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_augment_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python scripts/elem_microbench.py 64 2>&1 | tail -3
