#!/usr/bin/env python3
"""Where the warp instructions of a kernel go, by SOURCE LINE of the .cu file (ncu --import-source on + -lineinfo).
Prints the hottest source lines (warp instructions, share, threads per instruction, stall samples) and totals per
user-named line range.   usage: tools/ncu_source_regions.py report.ncu-rep [top_n] [lo-hi:label ...]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ranges = []
for a in sys.argv[3:]:
    r, label = a.split(":")
    lo, hi = r.split("-")
    ranges.append((int(lo), int(hi), label))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and "Instructions Executed" in r)
h = rows[hdr_i]
iL, iS = h.index("Line No") if "Line No" in h else 0, h.index("Source")
iW, iT, iSm = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
lines = []
for r in rows[hdr_i + 1:]:
    try:
        if r[2] != "-":
            continue
        lines.append((int(r[iL]), r[iS], float(r[iW] or 0), float(r[iT] or 0), float(r[iSm] or 0)))
    except (ValueError, IndexError):
        pass
tot = sum(l[2] for l in lines) or 1.0
tots = sum(l[4] for l in lines) or 1.0
print("total warp instructions %.0f, threads/inst %.2f, samples %.0f" % (tot, sum(l[3] for l in lines) / tot, tots))
for ln, src, w, t, sm in sorted(lines, key=lambda l: -l[2])[:top]:
    print("%5d %5.1f%% inst  %5.1f thr  %5.1f%% samples | %s" % (ln, 100 * w / tot, t / w if w else 0, 100 * sm / tots, src.strip()[:110]))
for lo, hi, label in ranges:
    sel = [l for l in lines if lo <= l[0] <= hi]
    w = sum(l[2] for l in sel)
    print("range %-28s lines %4d-%4d: %5.1f%% of warp instructions, %5.1f threads/inst, %5.1f%% of stall samples" %
          (label, lo, hi, 100 * w / tot, sum(l[3] for l in sel) / w if w else 0, 100 * sum(l[4] for l in sel) / tots))
