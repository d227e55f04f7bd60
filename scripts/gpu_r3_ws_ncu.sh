#!/bin/bash
# ncu --set full of the 64-output-channel halo kernel (weight-stationary MMAs, the default): L16 fprop 128->64 @512^2 and
# L1 fprop 64->64 @512^2.  Only the CSV pages travel back (the reports exceed the 64 MiB return limit).
mkdir -p gpurun_out /tmp/ncu
CMD="python scripts/conv_microbench.py --batch 64 --layers 16,1 --kinds fprop --reps 1"
$CMD > gpurun_out/r3_ws_ncu_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 1 -c 3 \
    -o /tmp/ncu/halo64 -f $CMD > gpurun_out/r3_ws_ncu.log 2>&1
ncu -i /tmp/ncu/halo64.ncu-rep --page raw --csv > gpurun_out/r3_halo64_raw.csv 2>/dev/null
ncu -i /tmp/ncu/halo64.ncu-rep --page source --csv > gpurun_out/r3_halo64_source.csv 2>/dev/null
tail -n 2 gpurun_out/r3_ws_ncu.log
ls -la gpurun_out/r3_halo64*
