#!/bin/bash
# A/B: vote-exit of the node / triangle loops (tune/libcrt_voteNL.so) against the in-tree while-while traversal.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CRT_LIB=$GRAFT_REPO_ROOT/tune/libcrt_batchpop.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or tight or replay" > gpurun_out/vote_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/vote_tests.log
tail -2 gpurun_out/vote_tests.log
: > gpurun_out/ab_vote.jsonl
run() { echo "{\"variant\": \"$1\"}" >> gpurun_out/ab_vote.jsonl; shift; timeout 600 python bench.py --no-cpu-baseline --warmup 3 "$@" >> gpurun_out/ab_vote.jsonl 2>> gpurun_out/ab_vote.err; }
for v in base batchpop base batchpop; do
  if [ $v = base ]; then unset CRT_LIB; else export CRT_LIB=$GRAFT_REPO_ROOT/tune/libcrt_$v.so; fi
  run $v --workload cornellSphere_1080p --steps 6
  run $v --workload monkeys_1080p --steps 4
  run $v --workload terrain1m_1080p --steps 2
done
python - <<'PY'
import json
v=None
for l in open("gpurun_out/ab_vote.jsonl"):
    try: j=json.loads(l)
    except Exception: continue
    if "variant" in j: v=j["variant"]; continue
    print(v, j["config"]["workload"], j["value"], {k: round(x) for k, x in j["stage_ms"].items()})
PY
