#!/bin/bash
# one `ncu --set full` capture of the bounce-1 traversal launch (extend 1 + shadow 0) per workload
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${TAG:-r02}
for WL in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_traverse --launch-skip 1 --launch-count 1 \
    -o gpurun_out/${TAG}_ncu_traverse_${WL} -f python tools/profile_step.py $WL 8 1 > gpurun_out/${TAG}_ncu_traverse_${WL}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
