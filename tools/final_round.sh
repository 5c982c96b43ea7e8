#!/bin/bash
# End-of-round evidence on one B200: full GPU test suite, the default bench line (what the driver runs), one bench line per
# workload, the reference arm, per-ray instruction counts, ncu launch list of bench.py itself, full captures of k_traverse
# and k_shade, compute-sanitizer.  Everything lands in gpurun_out/r02z_* and is copied to profiles/ by hand.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=gpurun_out/r02z
timeout 1800 python -m pytest tests -q -m gpu --durations=10 > ${T}_gpu_tests.log 2>&1
echo "tests rc=$?" >> ${T}_gpu_tests.log
tail -4 ${T}_gpu_tests.log
timeout 900 python tools/ncu_per_ray.py cornellSphere_1080p monkeys_1080p terrain1m_1080p cornell_4k > ${T}_per_ray.log 2>&1
cp gpurun_out/r02_traverse_per_ray.json profiles/r02_traverse_per_ray.json          # bench.py reads it from profiles/
cp gpurun_out/r02_traverse_per_ray.json ${T}_traverse_per_ray.json
timeout 900 python bench.py > ${T}_bench_default.json 2> ${T}_bench.err
: > ${T}_bench_all.jsonl
for wl in monkeys_1080p terrain1m_1080p cornell_4k cornell_1080p terrain131k_1080p cornell_256; do
  steps=4; [ "$wl" = terrain1m_1080p ] && steps=2
  timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-configs >> ${T}_bench_all.jsonl 2>> ${T}_bench.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${T}_bench_reference_arm.json 2>> ${T}_bench.err
: > ${T}_bench_stats.jsonl            # node visits / triangle tests per ray and lane statistics (profiling pass only)
for wl in cornellSphere_1080p monkeys_1080p terrain1m_1080p; do
  timeout 600 python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-configs --stats >> ${T}_bench_stats.jsonl 2>> ${T}_bench.err
done
python - <<'PY'
import json
j = json.load(open("gpurun_out/r02z_bench_default.json"))
print("default", j["value"], j["e2e"]["value"], j["seam_literal"]["value"], j["time_to_full_spp_s"], j["roofline"].get("frac"), j["roofline"].get("lanes_per_inst"), j["cuda_reference"], j["cpu_baseline"]["value"])
print({k: v["value"] for k, v in j["configs"].items()})
for l in open("gpurun_out/r02z_bench_all.jsonl"):
    j = json.loads(l)
    print(j["config"]["workload"], j["value"], j["mrays_per_s"], j["e2e"]["value"], j["seam_literal"]["value"], j["tight_bvh"], {k: round(v) for k, v in j["stage_ms"].items()}, j.get("cpu_baseline", {}).get("value"), j["roofline"].get("frac"))
for l in open("gpurun_out/r02z_bench_stats.jsonl"):
    j = json.loads(l); s = j["lane_stats"]
    print(j["config"]["workload"], j["tight_bvh"]["per_ray"], "node loop: %.1f lanes work, %.1f idle; triangle loop: %.1f work, %.1f idle; %.1f rays per refill" % (
        s["node_lanes"]/max(s["node_iters"],1), s["node_idle"]/max(s["node_iters"],1), s["leaf_lanes"]/max(s["leaf_iters"],1), s["leaf_idle"]/max(s["leaf_iters"],1), s["refill_rays"]/max(s["refills"],1)))
PY
cut -c1-300 ${T}_bench_reference_arm.json
tail -3 ${T}_bench.err
# ncu launch list of the bench command itself (first 600 launches)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${T}_launches_bench_py.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > ${T}_launches_bench_py.log 2>&1
# full captures: bounce-1 traversal launch and bounce-1 shade launch
for WL in cornellSphere_1080p terrain1m_1080p; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_traverse --launch-skip 1 --launch-count 1 \
    -o ${T}_ncu_traverse_${WL} -f python tools/profile_step.py $WL 8 1 > ${T}_ncu_traverse_${WL}.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_shade --launch-skip 1 --launch-count 1 \
  -o ${T}_ncu_shade_cornellSphere_1080p -f python tools/profile_step.py cornellSphere_1080p 8 1 > ${T}_ncu_shade_cornellSphere_1080p.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
bash tools/sanitize.sh
