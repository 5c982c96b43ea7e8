#!/bin/bash
# Whole-process wall time of the drop-in command line on BASELINE config 4 (999 700-triangle terrain, 1920x1080, 256 spp):
# OBJ load + scene assembly + crt_create + render + PNG, three runs each of the binaries built before (tune/old_cli, commit
# 56a8c83) and after the host-side preparation was threaded.  -> gpurun_out/r02zz_cli_time_to_image.json
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
A=/tmp/a707
python oracle/make_assets.py --out $A --terrain 707 > /dev/null 2>&1
mkdir -p $A/build
cd $A/build
python - > $GRAFT_REPO_ROOT/gpurun_out/r02zz_cli_time_to_image.json <<'PY'
import hashlib, json, os, subprocess, time
root = os.environ["GRAFT_REPO_ROOT"]
out = {"command": "./CudaRT -f terrain707 -s 256 -w 1920 -h 1080", "what": "wall_s = the whole process; render_s = the CLI's own 'Elapsed time' (launch loop only)"}
for tag, b in (("before_56a8c83", root + "/tune/old_cli/CudaRT"), ("after", root + "/cudart_b200/CudaRT")):
    runs = []
    for i in range(3):
        if os.path.exists("terrain707.png"):
            os.remove("terrain707.png")
        t = time.time()
        p = subprocess.run([b, "-f", "terrain707", "-s", "256", "-w", "1920", "-h", "1080"], capture_output=True, text=True)
        wall = time.time() - t
        el = [l for l in p.stdout.splitlines() if l.startswith("Elapsed time")]
        png = hashlib.sha1(open("terrain707.png", "rb").read()).hexdigest() if os.path.exists("terrain707.png") else None
        runs.append({"wall_s": round(wall, 3), "render_s": float(el[0].split("=")[1]) if el else None, "rc": p.returncode, "png_sha1": png,
                     "err": p.stderr[-200:] if p.returncode else None})
    out[tag] = runs
print(json.dumps(out))
PY
cat $GRAFT_REPO_ROOT/gpurun_out/r02zz_cli_time_to_image.json
