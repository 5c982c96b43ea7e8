#!/bin/bash
# full GPU check of the current tree: parity tests, then one bench line per workload
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
: > gpurun_out/bench_all.jsonl
timeout 900 python bench.py >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
for wl in monkeys_1080p cornell_1080p cornell_4k terrain131k_1080p terrain1m_1080p; do
  steps=4; [ "$wl" = terrain1m_1080p ] && steps=2
  timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
done
python - <<'PY'
import json
for l in open("gpurun_out/bench_all.jsonl"):
    j=json.loads(l)
    print(j["config"]["workload"], j["value"], j["e2e"]["value"], j["tight_bvh"], j["stage_ms"], j.get("cpu_baseline"), j["roofline"]["frac"])
PY
