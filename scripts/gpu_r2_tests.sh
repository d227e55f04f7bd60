#!/bin/bash
# round 2: new parity tests first (verbose, no -x so every failure shows), then the whole GPU suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_smi.txt 2>&1
python -m pytest tests/test_teacher_forced_gpu.py tests/test_envelope_gpu.py tests/test_repack_gpu.py -q -s -m gpu \
    > gpurun_out/r2_t_new.log 2>&1
echo "new tests rc=$?" >> gpurun_out/r2_t_new.log
python -m pytest tests/test_kernels_gpu.py -q -s -m gpu -k "full_size_conv_vs_cudnn" > gpurun_out/r2_t_fullsize.log 2>&1
echo "fullsize rc=$?" >> gpurun_out/r2_t_fullsize.log
python -m pytest tests -q -m gpu -x > gpurun_out/r2_t_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r2_t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
for f in gpurun_out/r2_t_new.log gpurun_out/r2_t_fullsize.log gpurun_out/r2_t_all.log gpurun_out/r2_smoke.log; do echo "== $f"; tail -n 5 $f; done
