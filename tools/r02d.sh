#!/bin/bash
# A/B: two wavefront batches in flight on two streams (crt_options.waves_in_flight) and traversal CTAs per SM
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r02d
: > $O.jsonl
for cfg in "1 0" "2 0" "2 3" "1 3" "2 2"; do
  set -- $cfg
  for wl in cornellSphere_1080p monkeys_1080p terrain1m_1080p; do
    steps=4; [ "$wl" = terrain1m_1080p ] && steps=2
    echo "{\"waves_in_flight\": $1, \"trav_blocks\": $2, \"workload\": \"$wl\"}" >> $O.jsonl
    CRT_WAVES_IN_FLIGHT=$1 CRT_TRAV_BLOCKS=$2 timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline >> $O.jsonl 2>> $O.err
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/r02d.jsonl"):
    j = json.loads(l)
    if "waves_in_flight" in j: print(j, end=" -> "); continue
    print(j["value"], j["e2e"]["value"], j["clocks"]["sm_mhz"])
PY
