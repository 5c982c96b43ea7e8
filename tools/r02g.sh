#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin.py tests/test_gpu_parity.py -q -m gpu -k "dropin or reference_main or progressive or command_line or literal_seam or display" 2>&1 | tail -8
timeout 1200 python bench.py > gpurun_out/r02g_bench_default.json 2> gpurun_out/r02g_bench_default.err
tail -3 gpurun_out/r02g_bench_default.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r02g_bench_default.json"))
for k in ("value", "ms_per_step", "mrays_per_s", "e2e", "seam_literal", "time_to_full_spp_s", "gpu_launches", "stage_ms", "roofline", "configs", "cuda_reference", "cpu_baseline", "clocks"):
    print(k, json.dumps(j.get(k))[:600])
PY
