mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"upsample2x_pad" -c 4 -o gpurun_out/prof_up -f python scripts/elem_microbench.py 64 > gpurun_out/ncu_up.log 2>&1
echo "ncu exit $?"
