#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r02e
: > $O.jsonl
for cfg in "1 0" "2 0" "2 3"; do
  set -- $cfg
  for wl in cornellSphere_1080p monkeys_1080p terrain1m_1080p; do
    steps=4; [ "$wl" = terrain1m_1080p ] && steps=2
    echo "{\"waves_in_flight\": $1, \"trav_blocks\": $2, \"workload\": \"$wl\"}" >> $O.jsonl
    CRT_TRAV_BLOCKS=$2 timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline --no-configs --waves-in-flight $1 >> $O.jsonl 2>> $O.err
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/r02e.jsonl"):
    j = json.loads(l)
    if "waves_in_flight" in j: print(j, end=" -> "); continue
    print(j["value"], j["e2e"]["value"], j["seam_literal"]["value"], j["time_to_full_spp_s"], j["stage_ms"], j["clocks"]["sm_mhz"])
PY
tail -5 $O.err
timeout 900 python tools/ncu_per_ray.py cornellSphere_1080p monkeys_1080p terrain1m_1080p 2>&1 | tail -5
