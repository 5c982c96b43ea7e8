#!/bin/bash
# Last GPU call of a round with a few GPU-minutes left: the driver's default bench line first (without the 20 s CPU baseline),
# then as much of the GPU test suite as fits.  Lands in gpurun_out/r02zz_*.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=gpurun_out/r02zz
timeout 80 python bench.py --no-cpu-baseline > ${T}_bench_default.json 2> ${T}_bench.err
echo "bench rc=$?"; cut -c1-400 ${T}_bench_default.json
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/r02zz_bench_default.json"))
    print("value", j["value"], "e2e", j["e2e"]["value"], "seam", j["seam_literal"]["value"], "create_s", j.get("create_s"))
    print({k: (v["value"], v.get("create_s")) for k, v in (j.get("configs") or {}).items()})
except Exception as e:
    print("no bench line:", e)
PY
timeout 150 python -m pytest tests -q -x -m gpu --durations=5 > ${T}_gpu_tests.log 2>&1
echo "tests rc=$?" >> ${T}_gpu_tests.log
tail -8 ${T}_gpu_tests.log
