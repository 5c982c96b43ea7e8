#!/bin/bash
# quick loop while changing kernels or the scheduler: the whole GPU suite, four bench lines, instructions per ray
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/${1:-check}
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
: > $O.jsonl
for wl in cornellSphere_1080p monkeys_1080p terrain1m_1080p cornell_4k; do
  steps=4; [ "$wl" = terrain1m_1080p ] && steps=2; [ "$wl" = cornell_256 ] && steps=64
  timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline --no-configs ${BENCH_ARGS} >> $O.jsonl 2>> $O.err
done
python - $O.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    j = json.loads(l)
    print(j["config"]["workload"], j["value"], j["e2e"]["value"], j["seam_literal"]["value"], {k: round(v) for k, v in j["stage_ms"].items()}, j["clocks"]["sm_mhz"])
PY
tail -5 $O.err
if [ -n "$PER_RAY" ]; then
  timeout 900 python tools/ncu_per_ray.py cornellSphere_1080p monkeys_1080p terrain1m_1080p 2>&1 | cut -c1-220 | tail -5
  cp gpurun_out/r02_traverse_per_ray.json $O.per_ray.json
fi
