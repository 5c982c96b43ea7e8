#!/bin/bash
# quick GPU check: the render-level parity tests, then bench lines for the two headline scenes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "replay or bitwise or counters or contract or sharding or rgba8" > gpurun_out/quick_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/quick_tests.log
tail -3 gpurun_out/quick_tests.log
: > gpurun_out/quick_bench.jsonl
for wl in cornellSphere_1080p monkeys_1080p ${EXTRA_WL}; do
  timeout 600 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline >> gpurun_out/quick_bench.jsonl 2>> gpurun_out/quick_bench.err
done
python - <<'PY'
import json
for l in open("gpurun_out/quick_bench.jsonl"):
    j=json.loads(l)
    print(j["config"]["workload"], j["value"], j["e2e"]["value"], j["stage_ms"])
PY
