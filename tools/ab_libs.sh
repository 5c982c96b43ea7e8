#!/bin/bash
# A/B of tuning builds: tools/ab_libs.sh <tag> <lib or "default"> ...   (three workloads each, one bench line per run)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; shift
O=gpurun_out/$TAG
: > $O.jsonl
for lib in "$@"; do
  for wl in cornellSphere_1080p monkeys_1080p terrain1m_1080p; do
    steps=4; [ "$wl" = terrain1m_1080p ] && steps=2
    echo "{\"lib\": \"$lib\", \"workload\": \"$wl\"}" >> $O.jsonl
    L=""; [ "$lib" != default ] && L=$PWD/$lib
    CRT_LIB=$L timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline --no-configs ${BENCH_ARGS} >> $O.jsonl 2>> $O.err
  done
done
python - $O.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    j = json.loads(l)
    if "lib" in j: print("%-28s %-20s" % (j["lib"], j["workload"]), end=" -> "); continue
    print(j["value"], j["seam_literal"]["value"], {k: round(v) for k, v in j["stage_ms"].items()})
PY
tail -3 $O.err
