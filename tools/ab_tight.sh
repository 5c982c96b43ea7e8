#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tight or golden" > gpurun_out/tight_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/tight_tests.log
: > gpurun_out/ab_tight.jsonl
for wl in cornellSphere_1080p monkeys_1080p cornell_1080p terrain131k_1080p terrain1m_1080p; do
  [ "$wl" = cornell_1080p ] && continue
  for t in on off auto; do
    steps=4; [ "$wl" = terrain1m_1080p ] && steps=2
    timeout 600 python bench.py --workload $wl --tight-bvh $t --steps $steps --warmup 3 --no-cpu-baseline >> gpurun_out/ab_tight.jsonl 2>> gpurun_out/ab_tight.err
    timeout 600 python bench.py --workload $wl --tight-bvh $t --steps 1 --warmup 3 --no-cpu-baseline --stats >> gpurun_out/ab_tight_stats.jsonl 2>> gpurun_out/ab_tight.err
  done
done
tail -3 gpurun_out/tight_tests.log
python - <<'PY'
import json
for f in ("gpurun_out/ab_tight.jsonl","gpurun_out/ab_tight_stats.jsonl"):
    for l in open(f):
        try: j=json.loads(l)
        except Exception: continue
        print(f[-20:], j["config"]["workload"], j["value"], j["e2e"]["value"], j["tight_bvh"], j["stage_ms"])
PY
