mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "upsample or bilinear or head" 2>&1 | tail -3
timeout 300 python scripts/elem_microbench.py 64 2>&1 | grep -i "upsample\|head"
