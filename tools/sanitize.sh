#!/bin/bash
# compute-sanitizer over a small render: static-pad scene (cornell) and dynamic-pad scene (8 194-triangle terrain)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for tool in memcheck racecheck; do
  : > gpurun_out/sanitizer_$tool.log
  for wl in cornell_256 terrain8k_256; do
    timeout 600 compute-sanitizer --tool $tool python tools/profile_step.py $wl 2 2 >> gpurun_out/sanitizer_$tool.log 2>&1
    # four waves in one call: two wave slots on two streams, accumulation chained by events
    CRT_ONECALL=1 timeout 600 compute-sanitizer --tool $tool python tools/profile_step.py $wl 1 4 >> gpurun_out/sanitizer_$tool.log 2>&1
  done
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sah" gpurun_out/sanitizer_$tool.log
done
