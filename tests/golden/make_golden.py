#!/usr/bin/env python3
"""tests/golden/make_golden.py - regenerate the golden fixtures from the UNMODIFIED reference.

Run in the dev container (needs /root/reference):

    make -C oracle ref oracle
    python oracle/make_assets.py
    python tests/golden/make_golden.py

For every fixture scene it runs oracle/_ref/ref_dump (the reference's own translation units, see
oracle/ref_dump.cpp) and stores under tests/golden/:

  <scene>.crts.gz      flattened scene arrays as the reference's Scene builds them (its loader, its
                       zero-init-pinned BVH builder, its light list) - input of every parity test
  <scene>.golden.npz   pretris  : Triangle[] before constructBvh (input of the product's BVH builder)
                       rays/hits: a seeded ray set and the records of the reference's intersectBVH
                       accum    : float radiance sums of the reference's samplePixel, srand(1)
                       rgba     : bytes through the reference's Renderer API (float3ToUchar4 + gamma)
                       accum_hi : the same sums for a 32x24 image at 256 spp (statistical parity test)
                       meta     : W, H, spp, seed

The ray set mixes (a) rays harvested from the path loop (camera, bounce and shadow rays),
(b) uniformly random rays, (c) adversarial rays: aimed exactly at vertices / edge midpoints (exact
t ties between neighbouring triangles), directions with exactly-zero components, the all-zero
direction produced by SURVEY.md D16, and non-default tMin/tMax windows.
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cudart_b200.scenefile import RAY_DTYPE, HIT_DTYPE, TRIANGLE_DTYPE, load_crts  # noqa: E402
from oracle import orc  # noqa: E402

ASSETS = os.path.join(ROOT, "oracle", "_ref", "assets")
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
OUT = os.path.dirname(os.path.abspath(__file__))

#        scene          W   H  spp
SCENES = [("cornell", 64, 48, 4), ("cornellSphere", 64, 48, 4), ("monkeys", 64, 48, 4), ("terrain64", 48, 32, 2),
          ("tiny", 40, 30, 4),          # tests/golden/tiny: hand-written loader corner cases, DOF camera, every BSDF
          ("soup", 48, 36, 4)]          # tests/golden/soup: oversized leaf, duplicate triangles, slivers, random soup
N_HARVEST, N_RANDOM, N_SPECIAL = 6000, 3000, 3000
HI_W, HI_H, HI_SPP = 32, 24, 256


def ref_dump(scene, W, H, *extra):
    assets = os.path.join(OUT, scene) if scene in ("tiny", "soup") else ASSETS
    cmd = [REF_DUMP, "--obj", f"{assets}/meshes/{scene}.obj", "--camera", f"{assets}/settings/{scene}-camera.json",
           "--textures", f"{assets}/textures/", "-w", str(W), "-h", str(H), *extra]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)


def make_rays(fs, rng):
    osc = orc.OracleScene(fs)
    _, log = osc.render_accum(1, rng=orc.RNG_COUNTER, seed=7, threads=1, ray_log_cap=1 << 20)
    pick = rng.choice(len(log), size=min(N_HARVEST, len(log)), replace=False)
    harvested = log[np.sort(pick)]

    lo = fs.nodes["min"][0].astype(np.float64)
    hi = fs.nodes["max"][0].astype(np.float64)
    ext = hi - lo
    rnd = np.zeros(N_RANDOM, RAY_DTYPE)
    rnd["o"] = (lo - 0.25 * ext + rng.random((N_RANDOM, 3)) * 1.5 * ext).astype(np.float32)
    d = rng.normal(size=(N_RANDOM, 3))
    rnd["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rnd["tmin"] = 1e-5
    rnd["tmax"] = np.finfo(np.float32).max

    sp = np.zeros(N_SPECIAL, RAY_DTYPE)
    sp["tmin"] = 1e-5
    sp["tmax"] = np.finfo(np.float32).max
    eye = fs.camera["eye"][0]
    tri = fs.tris[rng.integers(0, fs.n_tris, N_SPECIAL)]
    kind = rng.integers(0, 8, N_SPECIAL)
    v1, v2, v3 = tri["v1"], tri["v1"] + tri["e1"], tri["v1"] + tri["e2"]
    target = np.where((kind % 4 == 0)[:, None], v1, np.where((kind % 4 == 1)[:, None], v2,
                      np.where((kind % 4 == 2)[:, None], (v1 + v2) * np.float32(0.5), (v2 + v3) * np.float32(0.5))))
    origin = np.where((kind < 4)[:, None], eye[None, :],
                      (lo + rng.random((N_SPECIAL, 3)) * ext).astype(np.float32)).astype(np.float32)
    dirs = (target - origin).astype(np.float32)          # un-normalised on purpose for half of them
    nrm = np.linalg.norm(dirs, axis=1, keepdims=True).astype(np.float32)
    nrm[nrm == 0] = 1
    dirs = np.where((kind >= 2)[:, None], dirs / nrm, dirs).astype(np.float32)
    sp["o"], sp["d"] = origin, dirs
    # exact-zero direction components, -0, and the all-zero direction of D16
    z = rng.integers(0, N_SPECIAL, N_SPECIAL // 6)
    for k, idx in enumerate(z):
        m = k % 7
        if m == 6:
            sp["d"][idx] = 0
        else:
            sp["d"][idx, m % 3] = 0.0 if m < 3 else -0.0
    # origins exactly on node slab planes (0 * inf in the slab test)
    on = rng.integers(0, N_SPECIAL, N_SPECIAL // 10)
    nd = fs.nodes[rng.integers(0, fs.n_nodes, len(on))]
    ax = rng.integers(0, 3, len(on))
    sp["o"][on, ax] = nd["min"][np.arange(len(on)), ax]
    sp["d"][on, ax] = 0.0
    # tMin/tMax windows
    w = rng.integers(0, N_SPECIAL, N_SPECIAL // 8)
    sp["tmax"][w] = rng.random(len(w)).astype(np.float32) * np.float32(np.linalg.norm(ext))
    w = rng.integers(0, N_SPECIAL, N_SPECIAL // 16)
    sp["tmin"][w] = rng.random(len(w)).astype(np.float32) * np.float32(0.5 * np.linalg.norm(ext))
    return np.concatenate([harvested, rnd, sp])


def main():
    if not os.path.exists(REF_DUMP):
        sys.exit("oracle/_ref/ref_dump missing: run `make -C oracle ref` in the dev container first")
    rng = np.random.default_rng(20261017)
    for scene, W, H, spp in SCENES:
        with tempfile.TemporaryDirectory() as tmp:
            crts, pre = f"{tmp}/s.crts", f"{tmp}/pre.bin"
            ref_dump(scene, W, H, "--scene-out", crts, "--pretris-out", pre, "--positions-out", f"{tmp}/pos.bin",
                     "--premats-out", f"{tmp}/premats.bin")
            fs = load_crts(crts)
            rays = make_rays(fs, rng)
            rays.tofile(f"{tmp}/rays.bin")
            ref_dump(scene, W, H, "--rays-in", f"{tmp}/rays.bin", "--hits-out", f"{tmp}/hits.bin", "--spp", str(spp),
                     "--seed", "1", "--accum-out", f"{tmp}/accum.bin", "--rgba-out", f"{tmp}/rgba.bin")
            # a small, well-sampled image from the reference's samplePixel for the statistical parity test
            ref_dump(scene, HI_W, HI_H, "--spp", str(HI_SPP), "--seed", "1", "--accum-out", f"{tmp}/accum_hi.bin")
            accum_hi = np.fromfile(f"{tmp}/accum_hi.bin", np.float32).reshape(HI_H, HI_W, 3)
            hits = np.fromfile(f"{tmp}/hits.bin", HIT_DTYPE)
            accum = np.fromfile(f"{tmp}/accum.bin", np.float32).reshape(H, W, 3)
            rgba = np.fromfile(f"{tmp}/rgba.bin", np.uint8).reshape(H, W, 4)
            pretris = np.fromfile(pre, TRIANGLE_DTYPE)
            positions = np.fromfile(f"{tmp}/pos.bin", np.float32).reshape(-1, 9)
            premats = np.fromfile(f"{tmp}/premats.bin", np.float32).reshape(-1, 4, 3)
            with open(crts, "rb") as fi, gzip.GzipFile(f"{OUT}/{scene}.crts.gz", "wb", compresslevel=9, mtime=0) as fo:
                shutil.copyfileobj(fi, fo)
            np.savez_compressed(f"{OUT}/{scene}.golden.npz", pretris=pretris.view(np.uint8), rays=rays.view(np.uint8),
                                hits=hits.view(np.uint8), accum=accum, rgba=rgba, positions=positions, premats=premats,
                                accum_hi=accum_hi, meta=np.array([W, H, spp, 1, HI_W, HI_H, HI_SPP], np.int64))
            nh = int((hits["tri"] != 0xFFFFFFFF).sum())
            print(f"{scene}: {fs.n_tris} tris, {fs.n_nodes} nodes, {len(rays)} rays ({nh} hit), {W}x{H}x{spp}spp")


if __name__ == "__main__":
    main()
