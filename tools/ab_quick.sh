#!/bin/bash
# quick A/B of a tuning build (tune/libcrt_$1.so) against the in-tree library on the three traversal-heavy workloads
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$1
CRT_LIB=$GRAFT_REPO_ROOT/tune/libcrt_$V.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "replay or bitwise or counters or contract or sharding or rgba8" > gpurun_out/quick_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/quick_tests.log
tail -2 gpurun_out/quick_tests.log
: > gpurun_out/ab_quick.jsonl
run() { echo "{\"variant\": \"$1\"}" >> gpurun_out/ab_quick.jsonl; shift; timeout 600 python bench.py --no-cpu-baseline --warmup 3 "$@" >> gpurun_out/ab_quick.jsonl 2>> gpurun_out/ab_quick.err; }
for v in base $V; do
  if [ $v = base ]; then unset CRT_LIB; else export CRT_LIB=$GRAFT_REPO_ROOT/tune/libcrt_$v.so; fi
  run $v --workload cornellSphere_1080p --steps 6
  run $v --workload monkeys_1080p --steps 4
  run $v --workload cornell_1080p --steps 6
done
python - <<'PY'
import json
v=None
for l in open("gpurun_out/ab_quick.jsonl"):
    try: j=json.loads(l)
    except Exception: continue
    if "variant" in j: v=j["variant"]; continue
    print(v, j["config"]["workload"], j["value"], {k: round(x) for k, x in j["stage_ms"].items()})
PY
