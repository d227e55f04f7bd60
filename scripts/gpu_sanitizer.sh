#!/bin/bash
# compute-sanitizer over a tiny end-to-end exercise of every kernel family (race detection /
# sanitizers are new with this build; the reference has none).  Runs the target plain first.
# NOTE (round 2): on the graft GPU pool `compute-sanitizer` is closed by policy (it answers "closed on this pool", rc 86),
# so only the plain run of scripts/sanitizer_target.py executed there (all kernel families at tiny / ragged shapes: OK).
# The script is kept for pools where the tool is available.  In-kernel guards that replace it here: every mbarrier wait
# is bounded and traps instead of hanging (csrc/ptx.cuh mbar_wait), every launcher validates shapes / alignment / pitch
# before the launch (negative status -> RuntimeError naming kernel and shape), TMA clips ragged edges in hardware.
mkdir -p gpurun_out
timeout 300 python scripts/sanitizer_target.py > gpurun_out/r2s_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/r2s_plain.log; exit 1; }
for tool in memcheck synccheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python scripts/sanitizer_target.py \
      > gpurun_out/r2s_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r2s_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZER-TARGET-OK|Error:|hazard" gpurun_out/r2s_$tool.log | sort | uniq -c | head -n 12
done
