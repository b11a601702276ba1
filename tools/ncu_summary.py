#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i`): headline metrics, instruction mix and stall
reasons per barrier-delimited phase.  Usage: python tools/ncu_summary.py file.ncu-rep"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__shared_mem_per_block_dynamic',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print("---- kernel")
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print("  %-75s %s %s" % (k, r[i], units[i]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                      capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for si, st in enumerate(starts[:1]):
    h = rows[st]
    end = starts[si + 1] - 1 if si + 1 < len(starts) else len(rows)
    body = rows[st + 1:end]
    iS, iE, iSt = h.index('Source'), h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
    stallcols = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
    seg = 0
    segs = collections.defaultdict(lambda: [0, 0, 0])
    segstall = collections.defaultdict(collections.Counter)
    ops = collections.Counter()
    tot = 0
    for r in body:
        try:
            e, s = int(r[iE]), int(r[iSt])
        except Exception:
            continue
        if 'BAR.SYNC' in r[iS]:
            seg += 1
        segs[seg][0] += e
        segs[seg][1] += s
        segs[seg][2] += 1
        for c in stallcols:
            try:
                segstall[seg][c] += int(r[h.index(c)])
            except Exception:
                pass
        op = r[iS].split()
        op = (op[1] if op and op[0].startswith('@') and len(op) > 1 else (op[0] if op else '?')).split('.')[0]
        ops[op] += e
        tot += e
    print("SASS lines %d, warp instructions %d" % (len(body), tot))
    tots = sum(v[1] for v in segs.values())
    for k, v in segs.items():
        print("  phase %d: instr %5.1f%%  samples %5.1f%%  sass %5d  top stalls %s" % (
            k, 100 * v[0] / max(tot, 1), 100 * v[1] / max(tots, 1), v[2],
            [(a.replace('stall_', ''), b) for a, b in segstall[k].most_common(5)]))
    print("  mix:", ", ".join("%s %.1f%%" % (o, 100 * c / tot) for o, c in ops.most_common(14)))
