#!/bin/bash
# weak-scaling sweep on one 8-GPU box: the driver's launch line for N > 1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/scale.jsonl
timeout 300 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 4 --warmup 3 >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 4 --warmup 3 --workload cornell_4k >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
python - <<'PY'
import json
for l in open("gpurun_out/scale.jsonl"):
    try: j=json.loads(l)
    except Exception: continue
    print(j["config"]["workload"], j["n_gpus"], j["value"], j["e2e"]["value"], j["mrays_per_s"], j["clocks"])
PY
