#!/bin/bash
# bash tools/host_prepare_bench.sh [N=707] - phase times + hashes of the host scene preparation on the N x N terrain (CPU only).
# CSRC=<dir> builds against another checkout's csrc/ (before / after comparisons); REPS=<k> runs (default 6).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
N=${1:-707}
CSRC=${CSRC:-$ROOT/cudart_b200/csrc}
W=${TMPDIR:-/tmp}/crt_host_prepare_$N
mkdir -p $W
if [ ! -f $W/positions.bin ]; then
  PYTHONPATH=$ROOT python - $N $W <<'PY'
import sys, numpy as np
from cudart_b200 import procedural
tris, P = procedural.terrain_triangles(int(sys.argv[1]))
tris.tofile(sys.argv[2] + "/pre_tris.bin"); P.astype(np.float32).tofile(sys.argv[2] + "/positions.bin")
PY
fi
g++ -std=c++17 -O3 -ffp-contract=off -I$CSRC -o $W/bench $ROOT/tools/host_prepare_bench.cpp $CSRC/host_scene.cpp -lpthread
$W/bench $W
