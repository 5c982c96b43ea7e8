#!/bin/bash
# End-of-session evidence: full GPU test suite, one bench line per workload, the reference arm, ncu launch list and
# full captures of the traversal kernel on the headline scene and on the HBM-resident 999 700-triangle terrain.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
: > gpurun_out/bench_all.jsonl
timeout 900 python bench.py >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
for wl in monkeys_1080p cornell_1080p cornell_4k terrain131k_1080p terrain1m_1080p cornell_256; do
  steps=4; [ "$wl" = terrain1m_1080p ] && steps=2
  timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2>> gpurun_out/bench_all.err
: > gpurun_out/bench_stats.jsonl            # node visits / triangle tests per ray and lane statistics (slower: not bench values)
for wl in cornellSphere_1080p monkeys_1080p terrain1m_1080p; do
  timeout 600 python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --stats >> gpurun_out/bench_stats.jsonl 2>> gpurun_out/bench_all.err
done
python - <<'PY'
import json
for l in open("gpurun_out/bench_all.jsonl"):
    j=json.loads(l)
    print(j["config"]["workload"], j["value"], j["mrays_per_s"], j["e2e"]["value"], j["tight_bvh"], {k: round(v) for k, v in j["stage_ms"].items()}, j.get("cpu_baseline", {}).get("value"), j["roofline"]["frac"], j["clocks"])
PY
cat gpurun_out/bench_reference_arm.json | cut -c1-300
python - <<'PY'
import json
for l in open("gpurun_out/bench_stats.jsonl"):
    j=json.loads(l); s=j["lane_stats"]
    print(j["config"]["workload"], j["tight_bvh"]["per_ray"], "node loop: %.1f lanes work, %.1f idle; triangle loop: %.1f work, %.1f idle; %.1f rays per refill" % (
        s["node_lanes"]/max(s["node_iters"],1), s["node_idle"]/max(s["node_iters"],1), s["leaf_lanes"]/max(s["leaf_iters"],1), s["leaf_idle"]/max(s["leaf_iters"],1), s["refill_rays"]/max(s["refills"],1)))
PY
for WL in cornellSphere_1080p terrain1m_1080p; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${WL}.csv python tools/profile_step.py $WL 8 2 > gpurun_out/launches_${WL}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_traverse --launch-skip 1 --launch-count 1 \
    -o gpurun_out/ncu_traverse_${WL} -f python tools/profile_step.py $WL 8 1 > gpurun_out/ncu_traverse_${WL}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
