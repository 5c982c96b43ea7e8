#!/usr/bin/env python3
"""Upper bound on what ray re-ordering could buy the traversal kernel: time crt_trace_closest_device on the same
secondary/shadow rays in queue order, shuffled, and sorted by (direction octant, Morton code of the origin)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cudart_b200 import load_crts
from cudart_b200.renderer import B200Renderer
from oracle import orc
import multiprocessing as mp
from tools.parity_campaign import harvest_band

def morton3(q):
    def spread(v):
        v = v.astype(np.uint64) & 0x3ff
        v = (v | (v << 16)) & 0x30000ff
        v = (v | (v << 8)) & 0x300f00f
        v = (v | (v << 4)) & 0x30c30c3
        v = (v | (v << 2)) & 0x9249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)

name = sys.argv[1] if len(sys.argv) > 1 else "cornellSphere"
W, H, spp = 640, 480, 8
fs = load_crts(os.path.join(ROOT, "tests", "golden", name + ".crts.gz")).with_resolution(W, H)
with mp.Pool(os.cpu_count()) as pool:
    rays = np.concatenate(pool.map(harvest_band, [(name, W, H, spp, 5, r, min(r + 8, H)) for r in range(0, H, 8)]))
eye = fs.camera["eye"][0]
sec = rays[np.abs(rays["o"] - eye).max(axis=1) > 1e-4]          # drop camera rays
print(name, "rays", len(rays), "secondary+shadow", len(sec))
lo, hi = fs.nodes["min"][0], fs.nodes["max"][0]
q = np.clip((sec["o"] - lo) / np.maximum(hi - lo, 1e-9) * 1023, 0, 1023).astype(np.uint32)
octant = ((sec["d"][:, 0] < 0).astype(np.uint64) | ((sec["d"][:, 1] < 0).astype(np.uint64) << 1) | ((sec["d"][:, 2] < 0).astype(np.uint64) << 2))
key_full = (octant << 30) | morton3(q)
key_coarse = (octant << 9) | (morton3(q) >> 21)                   # 8 octants x 8x8x8 cells
m = morton3(q)
orders = {"queue order": np.arange(len(sec)), "shuffled": np.random.default_rng(0).permutation(len(sec)),
          "sorted octant+morton30": np.argsort(key_full, kind="stable"), "binned octant+8^3 cells": np.argsort(key_coarse, kind="stable"),
          "binned octant only (8)": np.argsort(octant, kind="stable"),
          "binned 4^3 cells (64)": np.argsort(m >> 24, kind="stable"),
          "binned octant+4^3 (512)": np.argsort((octant << 6) | (m >> 24), kind="stable"),
          "binned 8^3 cells (512)": np.argsort(m >> 21, kind="stable"),
          "binned octant+2^3 (64)": np.argsort((octant << 3) | (m >> 27), kind="stable"),
          "binned octant+16^3 (32k)": np.argsort((octant << 12) | (m >> 18), kind="stable")}
r = B200Renderer(fs, W, H)
hits = torch.empty(len(sec), 4, device="cuda")
for label, order in orders.items():
    d = torch.from_numpy(np.ascontiguousarray(sec[order]).view(np.float32).reshape(-1, 8)).cuda()
    ms = [r.trace_closest_device(d.data_ptr(), len(sec), hits.data_ptr()) for _ in range(5)]
    print("  %-26s %.3f ms  %.2f Grays/s" % (label, min(ms), len(sec) / min(ms) / 1e6))
r.close()
