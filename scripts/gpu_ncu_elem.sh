mkdir -p gpurun_out
timeout 300 python scripts/elem_microbench.py 64 > gpurun_out/plain_elem.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"upsample2x_pad|head1x1" -c 10 -o gpurun_out/prof_elem python scripts/elem_microbench.py 64 > gpurun_out/ncu_elem.log 2>&1
echo "ncu exit $?"
