#!/usr/bin/env python3
"""Per-ray instruction and DRAM figures of the traversal kernel, for bench.py's issue-slot roofline.

Runs `tools/profile_step.py <workload> <spp> 1` (one wavefront batch) under ncu for every k_traverse launch of the
wave, sums warp instructions, thread instructions and DRAM bytes over the launches and divides by the rays the wave
traced (the library's own counters).  Writes / updates profiles/r02_traverse_per_ray.json.  GPU box only.

usage: python tools/ncu_per_ray.py [workload ...]
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
SPP = {"cornellSphere_1080p": 8, "monkeys_1080p": 8, "terrain1m_1080p": 8, "cornell_4k": 2, "cornell_1080p": 8,
       "cornell_256": 16, "terrain131k_1080p": 8}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3,
        "nsecond": 1e-9, "second": 1.0}


def main():
    out_path = os.path.join(ROOT, "profiles", "r02_traverse_per_ray.json")
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for wl in sys.argv[1:] or ["cornellSphere_1080p", "monkeys_1080p", "terrain1m_1080p"]:
        spp = SPP.get(wl, 4)
        cmd = ["ncu", "--csv", "--clock-control", "none", "--kernel-name", "regex:k_traverse", "--metrics", ",".join(METRICS),
               sys.executable, os.path.join(ROOT, "tools", "profile_step.py"), wl, str(spp), "1"]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
        m = re.search(r"'extend_rays': (\d+), 'shadow_rays': (\d+)", p.stdout)
        rays = int(m.group(1)) + int(m.group(2))
        start = p.stdout.index('"ID"')
        rows = list(csv.DictReader(io.StringIO(p.stdout[start:])))
        tot = {k: 0.0 for k in METRICS}
        n_launch = set()
        for r in rows:
            k = r["Metric Name"]
            if k in tot:
                v = float(r["Metric Value"].replace(",", ""))
                tot[k] += v * UNIT.get(r["Metric Unit"], 1.0) if k != "smsp__issue_active.avg.pct_of_peak_sustained_active" else v
                n_launch.add(r["ID"])
        n = len(n_launch)
        table[wl] = {
            "warp_inst_per_ray": round(tot["smsp__inst_executed.sum"] / rays, 2),
            "thread_inst_per_ray": round(tot["smsp__thread_inst_executed.sum"] / rays, 2),
            "dram_bytes_per_ray": round((tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / rays, 2),
            "rays": rays, "launches": n, "spp_per_wave": spp,
            "kernel_seconds_under_ncu": round(tot["gpu__time_duration.sum"], 6),
            "issue_active_pct_mean_over_launches": round(tot["smsp__issue_active.avg.pct_of_peak_sustained_active"] / max(n, 1), 2),
            "source": "tools/ncu_per_ray.py: ncu --metrics %s over the %d k_traverse launches of one %d-spp wave" % (",".join(METRICS[:4]), n, spp),
        }
        print(wl, table[wl], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(table, open(os.path.join(ROOT, "gpurun_out", "r02_traverse_per_ray.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
