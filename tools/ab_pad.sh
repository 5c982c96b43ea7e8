#!/bin/bash
# A/B of the tight BVH's pad mode (baked into the boxes vs applied per node visit), parity tests and a small campaign.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tight or golden" > gpurun_out/pad_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/pad_tests.log
tail -3 gpurun_out/pad_tests.log
: > gpurun_out/ab_pad.jsonl
run() { echo "{\"variant\": \"$1\"}" >> gpurun_out/ab_pad.jsonl; shift; timeout 600 python bench.py --no-cpu-baseline --warmup 3 "$@" >> gpurun_out/ab_pad.jsonl 2>> gpurun_out/ab_pad.err; }
for pad in ${PADS:-dynamic}; do
  run $pad --workload terrain1m_1080p --steps 2 --pad $pad
  run $pad-stats --workload terrain1m_1080p --steps 1 --pad $pad --stats
  run $pad --workload terrain131k_1080p --steps 3 --pad $pad --tight-bvh on
  run $pad-stats --workload terrain131k_1080p --steps 1 --pad $pad --tight-bvh on --stats
  run $pad --workload monkeys_1080p --steps 4 --pad $pad
  run $pad-stats --workload monkeys_1080p --steps 1 --pad $pad --stats
  run $pad --workload cornellSphere_1080p --steps 6 --pad $pad
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
timeout 900 python tools/parity_campaign.py --rays ${RAYS:-3e6} --scenes cornell cornellSphere monkeys terrain64 soup tiny terrain256 terrain707 --out gpurun_out/parity_campaign_pad.json > gpurun_out/parity_campaign_pad.log 2>&1
echo "campaign rc=$?"; python - <<'PY'
import json
r=json.load(open("gpurun_out/parity_campaign_pad.json"))["results"]
for k,v in r.items(): print(k, v["rays"], {m: v[m] for m in v if m.endswith("mismatches")})
PY
