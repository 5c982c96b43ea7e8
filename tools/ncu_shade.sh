#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
WL=${1:-cornellSphere_1080p}
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_shade --launch-skip 1 --launch-count 1 \
  -o gpurun_out/ncu_shade_${WL} -f python tools/profile_step.py $WL 8 1 > gpurun_out/ncu_shade_${WL}.log 2>&1
ls -la gpurun_out/*.ncu-rep
