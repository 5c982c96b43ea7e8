#!/bin/bash
# Round 2, first GPU call: (1) baseline bench lines of the round-1 code on this box, (2) the same with the whole library
# built -fmad=false (what bit-exact shading would cost), (3) fresh `ncu --set full` captures of the FINAL k_shade (the
# round-1 capture predates the per-CTA queue slots), (4) the reference's own CUDA renderer on the same B200.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r02a
: > $O.bench.jsonl
for lib in "" tune/libcrt_fmadoff.so; do
  for wl in cornellSphere_1080p monkeys_1080p terrain1m_1080p; do
    steps=4; [ "$wl" = terrain1m_1080p ] && steps=2
    echo "{\"lib\": \"$lib\", \"workload\": \"$wl\"}" >> $O.bench.jsonl
    CRT_LIB=${lib:+$PWD/$lib} timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline >> $O.bench.jsonl 2>> $O.err
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/r02a.bench.jsonl"):
    j = json.loads(l)
    if "lib" in j: print(j); continue
    print(j["config"]["workload"], j["value"], j["mrays_per_s"], j["e2e"]["value"], {k: round(v) for k, v in j["stage_ms"].items()}, j["clocks"]["sm_mhz"])
PY
for WL in cornellSphere_1080p monkeys_1080p; do
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_shade --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r02a_ncu_shade_${WL} -f python tools/profile_step.py $WL 8 1 > gpurun_out/r02a_ncu_shade_${WL}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
# the reference's own CUDA renderer (parallel_renderer.cu:288-342), unmodified, on this GPU
A=$PWD/oracle/_ref/assets
printf "MESHES_PATH=%s/meshes\nCAMERA_PATH=%s/settings\nTEXTURES_PATH=%s/textures\n" $A $A $A > $A/.env
mkdir -p $A/build && cd $A/build
for sc in cornellSphere monkeys; do
  /usr/bin/time -f "wall %e s" env CUDA_VISIBLE_DEVICES=0 timeout 300 ../../CudaRT_ref -f $sc -s 64 -w 1920 -h 1080 2>&1 | tail -5
done
