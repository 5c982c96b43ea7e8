#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, on the CPU box): key metrics per kernel + SASS opcode mix.
usage: tools/ncu_summary.py report.ncu-rep [out.txt]"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum',
        'smsp__sass_average_branch_targets_threads_uniform.pct']
STALL = re.compile(r'smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio')


def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("== %s" % r[hdr.index('Kernel Name')], file=out)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("  %-78s %18s %s" % (k, r[i], units[i]), file=out)
    st = sorted(((float(r[i] or 0), STALL.match(h).group(1)) for i, h in enumerate(hdr) if STALL.match(h)), reverse=True)
    print("  stalls (warps per issue-active): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:7]), file=out)
src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv"]))))
starts = [i for i, r in enumerate(src) if r and r[0] == 'Kernel Name']
for n, k in enumerate(starts):
    h = src[k + 1]
    ci, ti, si = h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('Source')
    end = starts[n + 1] if n + 1 < len(starts) else len(src)
    body = [r for r in src[k + 2:end] if len(r) > ci and r[ci] != '']
    tot = sum(float(r[ci]) for r in body) or 1
    ops, thr = collections.Counter(), collections.Counter()
    for r in body:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si])
        op = m.group(2).split('.')[0] if m else '?'
        ops[op] += float(r[ci])
        thr[op] += float(r[ti])
    print("== SASS mix, launch %d: %d SASS lines, %.0f warp instructions, %.2f threads/instruction" %
          (n, len(body), tot, sum(thr.values()) / tot), file=out)
    print("  " + ", ".join("%s %.1f%%" % (op, 100 * c / tot) for op, c in ops.most_common(16)), file=out)
