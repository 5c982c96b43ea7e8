#!/bin/bash
# One 8-GPU box: the default bench at N = 1, 2, 4, 8 (what the driver's scaling run does: weak scaling of the headline
# workload, plus the config-5 frame under strong scaling in `configs`), and the explicit strong-scaling line of cornell_4k.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=gpurun_out/r02s
: > ${T}_scaling.jsonl
for n in 1 2 4 8; do
  timeout 900 python bench.py --gpus $n --steps 8 --warmup 3 --no-cpu-baseline >> ${T}_scaling.jsonl 2>> ${T}.err
done
: > ${T}_strong_4k.jsonl
for n in 1 2 4 8; do
  timeout 900 python bench.py --gpus $n --steps 4 --warmup 3 --workload cornell_4k --scaling strong --no-cpu-baseline --no-configs >> ${T}_strong_4k.jsonl 2>> ${T}.err
done
python - <<'PY'
import json
base = None
for l in open("gpurun_out/r02s_scaling.jsonl"):
    j = json.loads(l)
    base = base or j["value"]
    print("weak  N=%d value %.1f (%.1f %% of N x N=1) e2e %.1f seam %.1f ms/step %.2f" % (j["n_gpus"], j["value"], 100 * j["value"] / (base * j["n_gpus"]), j["e2e"]["value"], j["seam_literal"]["value"], j["ms_per_step"]),
          "| configs:", {k: (v["value"], v.get("scaling")) for k, v in (j.get("configs") or {}).items()}, "| cli:", j.get("cli_multi_gpu"))
base = None
for l in open("gpurun_out/r02s_strong_4k.jsonl"):
    j = json.loads(l)
    base = base or j["value"]
    print("strong cornell_4k N=%d value %.1f (%.1f %% of N x N=1) e2e %.1f ms/step %.2f time_to_full %s" % (j["n_gpus"], j["value"], 100 * j["value"] / (base * j["n_gpus"]), j["e2e"]["value"], j["ms_per_step"], j.get("time_to_full_spp_s")))
PY
tail -5 ${T}.err
