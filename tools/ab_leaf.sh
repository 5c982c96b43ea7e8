#!/bin/bash
# A/B of the traversal's triangle test: tune/libcrt_old.so (early exits, IEEE divide) against the in-tree library
# (exit-free Moeller-Trumbore, short reciprocal), plus lane statistics of the in-tree kernel.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "reciprocal or rebuilds or golden or tight or replay or update_scene" > gpurun_out/ab_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ab_tests.log
tail -3 gpurun_out/ab_tests.log
: > gpurun_out/ab_leaf.jsonl
for wl in cornellSphere_1080p monkeys_1080p; do
  for v in old new old new; do
    lib=$GRAFT_REPO_ROOT/tune/libcrt_$v.so; [ "$v" = new ] && lib=$GRAFT_REPO_ROOT/cudart_b200/libcrt_b200.so
    echo "{\"variant\": \"$v\"}" >> gpurun_out/ab_leaf.jsonl
    CRT_LIB=$lib timeout 600 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline >> gpurun_out/ab_leaf.jsonl 2>> gpurun_out/ab_leaf.err
  done
done
for wl in cornellSphere_1080p monkeys_1080p terrain1m_1080p; do
  echo "{\"variant\": \"stats\"}" >> gpurun_out/ab_leaf.jsonl
  timeout 600 python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --stats >> gpurun_out/ab_leaf.jsonl 2>> gpurun_out/ab_leaf.err
done
python - <<'PY'
import json
v=None
for l in open("gpurun_out/ab_leaf.jsonl"):
    try: j=json.loads(l)
    except Exception: continue
    if "variant" in j: v=j["variant"]; continue
    print(v, j["config"]["workload"], j["value"], j["e2e"]["value"], j["stage_ms"], j.get("lane_stats"), j["tight_bvh"]["per_ray"])
PY
