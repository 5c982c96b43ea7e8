#!/bin/bash
# Where the drop-in command line's time to the image goes on BASELINE config 4 (CRT_CLI_TIMING=1), and that the frame it
# writes holds the same pixels as the one the pre-threading build (tune/old_cli, when present) writes.
# -> gpurun_out/r02zz_cli_phase_timing.txt
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
A=/tmp/a707
python oracle/make_assets.py --out $A --terrain 707 > /dev/null 2>&1
mkdir -p $A/build $A/old
O=$GRAFT_REPO_ROOT/gpurun_out/r02zz_cli_phase_timing.txt
echo "# CRT_CLI_TIMING=1 ./CudaRT -f terrain707 -s 256 -w 1920 -h 1080 (999 700 triangles); nproc=$(nproc)" > $O
cd $A/build
for i in 1 2; do
  s=$(date +%s.%N)
  CRT_CLI_TIMING=1 $GRAFT_REPO_ROOT/cudart_b200/CudaRT -f terrain707 -s 256 -w 1920 -h 1080 2>&1 | grep -E "timing|Elapsed" >> $O
  e=$(date +%s.%N)
  python -c "print('whole process %.3f s' % ($e - $s))" >> $O
done
if [ -x $GRAFT_REPO_ROOT/tune/old_cli/CudaRT ]; then
  cd $A/old
  $GRAFT_REPO_ROOT/tune/old_cli/CudaRT -f terrain707 -s 256 -w 1920 -h 1080 > /dev/null 2>&1
  python - $A/build/terrain707.png $A/old/terrain707.png >> $O <<'PY'
import struct, sys, zlib, hashlib
def pixels(path):
    data = open(path, "rb").read()
    off, idat = 8, b""
    while off < len(data):
        n, tag = struct.unpack(">I4s", data[off:off + 8])
        if tag == b"IDAT":
            idat += data[off + 8:off + 8 + n]
        off += 12 + n
    return hashlib.sha1(zlib.decompress(idat)).hexdigest(), len(data)
a, b = pixels(sys.argv[1]), pixels(sys.argv[2])
print("decoded frame sha1 (banded PNG writer) %s, %d bytes; (pre-threading build) %s, %d bytes; identical pixels: %s" % (a[0], a[1], b[0], b[1], a[0] == b[0]))
PY
fi
cat $O
