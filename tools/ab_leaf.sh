#!/bin/bash
# A/B of the traversal's leaf phase: tune/libcrt_old.so (per-leaf for loop, early exits), tune/libcrt_flatbranchy.so
# (flat loop, early exits) and the in-tree library (flat loop, exit-free Moeller-Trumbore, short reciprocal).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "reciprocal or rebuilds or golden or tight or replay or update_scene" > gpurun_out/ab_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ab_tests.log
tail -3 gpurun_out/ab_tests.log
: > gpurun_out/ab_leaf.jsonl
for wl in cornellSphere_1080p monkeys_1080p terrain1m_1080p; do
  for v in old flatbranchy new; do
    [ "$wl" = terrain1m_1080p ] && [ "$v" = flatbranchy ] && continue
    steps=6; [ "$wl" = terrain1m_1080p ] && steps=2
    lib=$GRAFT_REPO_ROOT/tune/libcrt_$v.so; [ "$v" = new ] && lib=$GRAFT_REPO_ROOT/cudart_b200/libcrt_b200.so
    echo "{\"variant\": \"$v\"}" >> gpurun_out/ab_leaf.jsonl
    CRT_LIB=$lib timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline >> gpurun_out/ab_leaf.jsonl 2>> gpurun_out/ab_leaf.err
  done
done
python - <<'PY'
import json
v=None
for l in open("gpurun_out/ab_leaf.jsonl"):
    try: j=json.loads(l)
    except Exception: continue
    if "variant" in j: v=j["variant"]; continue
    print(v, j["config"]["workload"], j["value"], j["e2e"]["value"], j["e2e"].get("h2d_bytes_per_step"), j["e2e"].get("host_rebuilds"), j["stage_ms"])
PY
# per-SASS-line instruction counts of the new traversal kernel (bounce-1 launch)
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_traverse --launch-skip 1 --launch-count 1 \
  -o gpurun_out/ncu_traverse_v2 -f python tools/profile_step.py cornellSphere_1080p 8 1 > gpurun_out/ncu_traverse_v2.log 2>&1
ls -la gpurun_out/*.ncu-rep
