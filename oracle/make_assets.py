#!/usr/bin/env python3
"""oracle/make_assets.py - TEST INFRASTRUCTURE: author the scene fixtures both sides render.

None of the reference's three bundled scenes runs as shipped: the camera
JSONs of cornell / cornellSphere are missing, their .mtl files have no `type EMISSIVE` (zero
lights -> NaN), texture-less scenes crash the texture loader, and monkeys.mtl names two PNGs that
are not in the tree.  This script writes a self-contained asset directory

    <out>/meshes/{cornell,cornellSphere,monkeys,terrainN}.{obj,mtl}
    <out>/settings/<scene>-camera.json
    <out>/textures/{marble,wood}.png
    <out>/.env

(default <out> = oracle/_ref/assets, git-ignored, shipped to the GPU box by gpurun).  The
bundled .obj files are copied from /root/reference/meshes when that tree is present (they are
inputs, fed unchanged to BOTH the reference and the new path); everything else is authored here
and documented as a deviation from the shipped assets:

  * `type EMISSIVE` appended to lambert4SG (the Ka 1 1 1 ceiling quad) in cornell / cornellSphere;
  * one unused material with a map_Kd so TextureStore::loadAll sees >= 1 texture (D14);
  * cornellSphere (BASELINE config 2, "specular+refractive materials"): the sphere's lambert5SG
    becomes REFRACTIVE (Ks 1 1 1, Ni 1.5), the tall box (pCube3) gets a new SPECULAR material
    and the short box (pCube2) a new DIFFSPEC material, by re-pointing their `usemtl` lines;
  * two procedural 64x64 RGB PNGs;
  * camera JSONs for cornell / cornellSphere (same keys as settings/monkeys-camera.json);
  * `terrainN`: the procedural height-field scene of BASELINE config 4 (N x N quads).
"""
import argparse
import json
import math
import os
import shutil
import struct
import sys
import zlib

REF = "/root/reference"

CORNELL_CAMERA = {
    "depthOfField": 0, "fStop": 1000.0, "eye": [2.6, 0.548, 0.0], "name": "main", "focalLength": 0.032,
    "fieldOfView": 60.0, "focusDistance": 2.6, "upDirection": [0.0, 1.0, 0.0], "viewDirection": [-2.6, 0.0, 0.0],
}

DUMMY_TEX_MTL = """newmtl zzDummyTex
illum 4
Kd 0.50 0.50 0.50
Ka 0.00 0.00 0.00
Tf 1.00 1.00 1.00
Ni 1.00
map_Kd wood.png
type LAMBERT
"""


def write_png(path, w, h, rgb_fn):
    raw = bytearray()
    for y in range(h):
        raw.append(0)
        for x in range(w):
            raw.extend(bytes(rgb_fn(x, y)))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
    png += chunk(b"IDAT", zlib.compress(bytes(raw), 9)) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)


def marble(x, y):
    v = 0.5 + 0.5 * math.sin(0.35 * x + 2.5 * math.sin(0.21 * y) + 1.3 * math.sin(0.13 * (x + y)))
    g = int(140 + 100 * v)
    return (g, g, min(255, g + 10))


def wood(x, y):
    r = math.hypot(x - 20.0, y - 70.0)
    v = 0.5 + 0.5 * math.sin(1.1 * r + 0.8 * math.sin(0.3 * x))
    return (int(120 + 70 * v), int(70 + 45 * v), int(30 + 25 * v))


def patch_mtl(text, emissive=("lambert4SG",), overrides=None, extra=""):
    """Append keys to named material blocks of an MTL text."""
    overrides = overrides or {}
    out, cur = [], None

    def flush():
        if cur in emissive:
            out.append("type EMISSIVE")
        for line in overrides.get(cur, []):
            out.append(line)

    for line in text.splitlines():
        if line.startswith("newmtl "):
            flush()
            cur = line.split()[1]
        out.append(line)
    flush()
    return "\n".join(out) + "\n" + extra


def terrain_obj(path_obj, path_mtl, n, extent=30.0):
    """N x N quad height field (2 N^2 triangles) + a 12x12 emissive quad above it.

    y = 1.5 sin(0.35x) cos(0.35z) + 0.25 sin(3.1x + 1.7z); triangles have area >~ 1e-3 so none is
    dropped by the reference's absolute |det| < 1e-5 test (D9).  One `usemtl` per `g` (D25).
    """
    name = os.path.splitext(os.path.basename(path_obj))[0]
    with open(path_mtl, "w") as f:
        f.write("newmtl ground\nillum 4\nKd 0.55 0.50 0.42\nKa 0.00 0.00 0.00\nNi 1.00\ntype LAMBERT\n")
        f.write("newmtl lamp\nillum 4\nKd 1.00 1.00 1.00\nKa 1.00 0.95 0.85\nNi 1.00\ntype EMISSIVE\n")
        f.write(DUMMY_TEX_MTL)
    with open(path_obj, "w") as f:
        f.write("mtllib %s.mtl\ng default\n" % name)
        step = 2.0 * extent / n
        for j in range(n + 1):
            z = -extent + j * step
            rows = []
            for i in range(n + 1):
                x = -extent + i * step
                y = 1.5 * math.sin(0.35 * x) * math.cos(0.35 * z) + 0.25 * math.sin(3.1 * x + 1.7 * z)
                rows.append("v %.6f %.6f %.6f\n" % (x, y, z))
            f.write("".join(rows))
        f.write("vt 0.000000 0.000000\nvn 0.000000 1.000000 0.000000\n")
        f.write("g terrain\nusemtl ground\n")
        for j in range(n):
            rows = []
            for i in range(n):
                a = j * (n + 1) + i + 1
                b = a + 1
                c = a + (n + 1)
                d = c + 1
                rows.append("f %d/1/1 %d/1/1 %d/1/1\nf %d/1/1 %d/1/1 %d/1/1\n" % (a, c, b, b, c, d))
            f.write("".join(rows))
        base = (n + 1) * (n + 1)
        f.write("g default\n")
        for (x, z) in ((-6, -6), (6, -6), (-6, 6), (6, 6)):
            f.write("v %.6f 14.000000 %.6f\n" % (x, z))
        f.write("vn 0.000000 -1.000000 0.000000\n")
        f.write("g light\nusemtl lamp\n")
        f.write("f %d/1/2 %d/1/2 %d/1/2\nf %d/1/2 %d/1/2 %d/1/2\n"
                % (base + 1, base + 2, base + 3, base + 3, base + 2, base + 4))


TERRAIN_CAMERA = {
    "depthOfField": 0, "fStop": 1000.0, "eye": [0.0, 12.0, 40.0], "name": "main", "focalLength": 0.032,
    "fieldOfView": 60.0, "focusDistance": 40.0, "upDirection": [0.0, 1.0, 0.0],
    "viewDirection": [0.0, -13.68, -37.59],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "assets"))
    ap.add_argument("--terrain", type=int, nargs="*", default=[64], help="height-field sizes N to generate (707 -> 1M tris)")
    args = ap.parse_args()
    out = args.out
    for d in ("meshes", "settings", "textures"):
        os.makedirs(os.path.join(out, d), exist_ok=True)
    with open(os.path.join(out, ".env"), "w") as f:
        f.write("MESHES_PATH=%s/meshes\nCAMERA_PATH=%s/settings\nTEXTURES_PATH=%s/textures\n" % (out, out, out))

    write_png(os.path.join(out, "textures", "marble.png"), 64, 64, marble)
    write_png(os.path.join(out, "textures", "wood.png"), 64, 64, wood)

    src = os.path.join(REF, "meshes")
    if os.path.isdir(src):
        # cornell: lights + dummy texture
        shutil.copy(os.path.join(src, "cornell.obj"), os.path.join(out, "meshes", "cornell.obj"))
        mtl = open(os.path.join(src, "cornell.mtl")).read()
        open(os.path.join(out, "meshes", "cornell.mtl"), "w").write(patch_mtl(mtl, extra=DUMMY_TEX_MTL))
        # cornellSphere: refractive sphere, mirror tall box, glossy short box
        obj = open(os.path.join(src, "cornellSphere.obj")).read().splitlines()
        for i, line in enumerate(obj):
            if line.startswith("g group1 pCube3") and obj[i + 1].startswith("usemtl"):
                obj[i + 1] = "usemtl mirrorSG"
            if line.startswith("g group1 pCube2") and obj[i + 1].startswith("usemtl"):
                obj[i + 1] = "usemtl glossySG"
        open(os.path.join(out, "meshes", "cornellSphere.obj"), "w").write("\n".join(obj) + "\n")
        mtl = open(os.path.join(src, "cornellSphere.mtl")).read()
        extra = ("newmtl mirrorSG\nillum 4\nKd 0.00 0.00 0.00\nKa 0.00 0.00 0.00\nNi 1.00\nKs 0.90 0.90 0.90\ntype SPECULAR\n"
                 "newmtl glossySG\nillum 4\nKd 0.70 0.65 0.20\nKa 0.00 0.00 0.00\nNi 1.00\nKs 0.80 0.80 0.80\n"
                 "type DIFFSPEC\ndiffuse 0.6\n" + DUMMY_TEX_MTL)
        over = {"lambert5SG": ["Ks 1.00 1.00 1.00", "Ni 1.50", "type REFRACTIVE"]}
        text = patch_mtl(mtl, overrides=over, extra=extra)
        # the shipped block has `Ni 1.00` before our override; the loader keeps the last value
        open(os.path.join(out, "meshes", "cornellSphere.mtl"), "w").write(text)
        for n in ("monkeys.obj", "monkeys.mtl"):
            shutil.copy(os.path.join(src, n), os.path.join(out, "meshes", n))
        shutil.copy(os.path.join(REF, "settings", "monkeys-camera.json"), os.path.join(out, "settings", "monkeys-camera.json"))
    else:
        print("make_assets: %s absent - bundled meshes not refreshed" % src, file=sys.stderr)

    for name in ("cornell", "cornellSphere"):
        json.dump(CORNELL_CAMERA, open(os.path.join(out, "settings", name + "-camera.json"), "w"))

    for n in args.terrain:
        name = "terrain%d" % n
        terrain_obj(os.path.join(out, "meshes", name + ".obj"), os.path.join(out, "meshes", name + ".mtl"), n)
        json.dump(TERRAIN_CAMERA, open(os.path.join(out, "settings", name + "-camera.json"), "w"))
    print("assets written to", out)


if __name__ == "__main__":
    main()
