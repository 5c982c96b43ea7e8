#!/bin/bash
# A/B sweep of the tight-BVH builder knobs (temporary env overrides), tight BVH forced on
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/sweep_tight.jsonl
for cfg in "1.5 8" "2 8" "2 6" "2.7 6" "3.5 8" "2 12" "3.5 12" "1 6"; do
  set -- $cfg
  for wl in cornellSphere_1080p monkeys_1080p terrain1m_1080p; do
    steps=3; [ "$wl" = terrain1m_1080p ] && steps=2
    CRT_TRAV=$1 CRT_MAXLEAF=$2 timeout 600 python bench.py --workload $wl --tight-bvh on --steps $steps --warmup 3 --no-cpu-baseline --stats 2>> gpurun_out/sweep.err | \
      python -c "
import sys, json
j=json.loads(sys.stdin.read()); j['cfg']='$cfg'; print(json.dumps(j))" >> gpurun_out/sweep_tight.jsonl
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/sweep_tight.jsonl"):
    j=json.loads(l)
    print(j["cfg"], j["config"]["workload"], j["value"], j["tight_bvh"], j["stage_ms"]["extend"])
PY
