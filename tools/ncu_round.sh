#!/bin/bash
# ncu evidence for the current tree: launch list with traffic metrics, and one full capture of a traversal launch
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
WL=${1:-cornellSphere_1080p}
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_${WL}.csv python tools/profile_step.py $WL 8 2 > gpurun_out/launches_${WL}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_traverse --launch-skip 1 --launch-count 1 \
  -o gpurun_out/ncu_traverse_${WL} -f python tools/profile_step.py $WL 8 1 > gpurun_out/ncu_traverse_${WL}.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/launches_${WL}.log
