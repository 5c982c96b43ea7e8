#!/bin/bash
# A/B: pad mode of the tight BVH as chosen per scene, and the refill threshold of the traversal (tune/libcrt_refillN.so).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/ab_pad.jsonl
run() { echo "{\"variant\": \"$1\"}" >> gpurun_out/ab_pad.jsonl; shift; timeout 600 python bench.py --no-cpu-baseline --warmup 3 "$@" >> gpurun_out/ab_pad.jsonl 2>> gpurun_out/ab_pad.err; }
run auto --workload monkeys_1080p --steps 4
run auto --workload terrain1m_1080p --steps 2
run auto --workload cornellSphere_1080p --steps 6
run auto --workload cornell_1080p --steps 6
run dynamic --workload cornell_1080p --steps 6 --pad dynamic
for v in refill8 refill24; do
  export CRT_LIB=$GRAFT_REPO_ROOT/tune/libcrt_$v.so
  run $v --workload monkeys_1080p --steps 4
  run $v --workload terrain1m_1080p --steps 2
  run $v --workload cornellSphere_1080p --steps 6
  unset CRT_LIB
done
python - <<'PY'
import json
v=None
for l in open("gpurun_out/ab_pad.jsonl"):
    try: j=json.loads(l)
    except Exception: continue
    if "variant" in j: v=j["variant"]; continue
    print(v, j["config"]["workload"], j["value"], j["e2e"]["value"], j["tight_bvh"], {k: round(x) for k, x in j["stage_ms"].items()})
PY
