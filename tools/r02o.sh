#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r02o
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "closest_hit or bitwise or replay_is_bit_exact" 2>&1 | tail -3
: > $O.jsonl
for spw in 0 8 16; do
  for wl in cornellSphere_1080p monkeys_1080p terrain1m_1080p; do
    steps=4; [ "$wl" = terrain1m_1080p ] && steps=2
    echo "{\"spp_per_wave\": $spw, \"workload\": \"$wl\"}" >> $O.jsonl
    timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline --no-configs --spp-per-wave $spw >> $O.jsonl 2>> $O.err
  done
done
python - $O.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    j = json.loads(l)
    if "spp_per_wave" in j and "workload" in j and "value" not in j: print(j, end=" -> "); continue
    print(j["value"], j["e2e"]["value"], j["seam_literal"]["value"], j["config"]["spp_per_wave"])
PY
tail -3 $O.err
