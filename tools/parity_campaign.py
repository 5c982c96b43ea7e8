#!/usr/bin/env python3
"""Large bit-exactness campaign for the closest-hit contract (">= 1e7 rays per scene including
harvested secondary and shadow rays").  For every fixture scene:

  1. harvest every ray the path loop hands to the traversal (camera, bounce and shadow rays) from the CPU oracle,
     in parallel over row bands;
  2. trace them with the oracle's restatement of the reference traversal (pinned bit-exact to the reference);
  3. trace them on the B200 through the C ABI: the render path's traversal with the tight per-triangle BVH forced on (pad baked into
     its boxes, and tight boxes padded per node visit) and forced off (distance-culled all three), and exact-visit (no culling, reference leaves only);
  4. compare (triangle index, t, u, v) bit for bit.

usage (GPU box): python tools/parity_campaign.py [--rays 1e7] [--out profiles/r01_parity_campaign.json]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def scene(name, W, H):
    """Fixture scenes by name; terrainN (N != 64) is generated and built by the product (bench.load_scene)."""
    import bench
    return bench.load_scene(name, W, H)


def harvest_band(args):
    name, W, H, spp, seed, r0, r1 = args
    from oracle import orc
    fs = scene(name, W, H)
    _, log = orc.OracleScene(fs).render_accum(spp, rng=orc.RNG_COUNTER, seed=seed, threads=1, rows=(r0, r1),
                                              ray_log_cap=(r1 - r0) * W * spp * 12)
    return log


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=float, default=1e7)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_campaign.json"))
    ap.add_argument("--scenes", nargs="*", default=["cornell", "cornellSphere", "monkeys", "terrain64", "soup", "tiny"])
    a = ap.parse_args()
    from cudart_b200 import load_crts
    from cudart_b200.renderer import B200Renderer
    from oracle import orc

    W, H = 640, 480
    results = {}
    workers = os.cpu_count() or 1
    for name in a.scenes:
        fs = scene(name, W, H)
        rays_wanted = a.rays if len(fs.tris) < 100_000 else a.rays / 8      # the oracle walks ~2000 nodes per ray on 1M triangles
        spp = max(1, int(rays_wanted / (W * H * 3.0)))           # ~3-6 rays per sample
        t0 = time.time()
        bands = [(name, W, H, spp, 2026, r, min(r + 8, H)) for r in range(0, H, 8)]
        with mp.Pool(workers if len(fs.tris) < 100_000 else min(workers, 8)) as pool:
            rays = np.concatenate(pool.map(harvest_band, bands))
        t1 = time.time()
        want = orc.OracleScene(fs).trace_closest(rays)
        t2 = time.time()
        rec = {"rays": int(len(rays)), "hits": int((want["tri"] != 0xFFFFFFFF).sum()), "harvest_s": round(t1 - t0, 1),
               "oracle_trace_s": round(t2 - t1, 1)}
        for mode, exact, tight, dyn in (("culled_tight_bvh", False, True, False), ("culled_tight_bvh_dynamic_pad", False, True, True),
                                        ("culled", False, False, None), ("exact_visit", True, False, None)):
            r = B200Renderer(fs, W, H, tight_bvh=tight, dynamic_pad=dyn)
            if tight:
                rec[mode + "_wide_nodes"], rec[mode + "_split_triangles"] = r.tight_info()
            got = r.trace_closest(rays, exact_visit=exact)
            r.close()
            bad = (got["tri"] != want["tri"]) | (got["t"].view(np.uint32) != want["t"].view(np.uint32)) | \
                  (got["u"].view(np.uint32) != want["u"].view(np.uint32)) | (got["v"].view(np.uint32) != want["v"].view(np.uint32))
            rec[mode + "_mismatches"] = int(bad.sum())
            if bad.any():
                k = int(np.nonzero(bad)[0][0])
                rec[mode + "_first"] = {"ray": rays[k].tolist(), "got": got[k].tolist(), "want": want[k].tolist()}
        results[name] = rec
        print(name, rec, flush=True)
    json.dump({"what": "closest-hit records of the B200 path vs the oracle restatement of the reference traversal, bit for bit",
               "image_for_harvest": [W, H], "results": results}, open(a.out, "w"), indent=1)
    bad = sum(v["culled_tight_bvh_mismatches"] + v["culled_tight_bvh_dynamic_pad_mismatches"] + v["culled_mismatches"] +
              v["exact_visit_mismatches"] for v in results.values())
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
