#!/usr/bin/env python
"""Per-source-line summary of ONE kernel of an ncu report: executed warp instructions, FP64
arithmetic instructions and stall samples per line of our files, plus the opcode mix.
Needs -lineinfo + --import-source on.
Usage: ncu_kernel_lines.py report.ncu-rep <substring of the kernel name> [top-N]"""
import collections
import csv
import io
import subprocess
import sys

rep, want = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
FP64 = ("DADD", "DMUL", "DFMA")
per_line = collections.defaultdict(lambda: [0, 0, 0])  # inst, fp64 inst, samples
ops = collections.defaultdict(int)
total = [0, 0, 0]
cur_file, func, hdr, key = None, None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1]
        key = None
        continue
    if r[0] == "Line No":
        hdr = r
        i_line, i_src, i_addr, i_sass = 0, 1, 2, 3
        i_n, i_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or func is None or want not in func or len(r) <= max(i_n, i_s):
        continue
    if r[i_addr] == "-":  # a source line; its SASS rows follow
        key = (cur_file, r[i_line], r[i_src].strip()[:105])
        continue
    if not r[i_addr].startswith("0x") or key is None:
        continue
    try:
        n, s = int(r[i_n]), int(r[i_s])
    except ValueError:
        continue
    w = r[i_sass].split()
    o = w[1] if (w and w[0].startswith("@") and len(w) > 1) else (w[0] if w else "?")
    o = o.split(".")[0]
    ops[o] += n
    per_line[key][0] += n
    per_line[key][2] += s
    total[0] += n
    total[2] += s
    if o in FP64:
        per_line[key][1] += n
        total[1] += n
print("kernel matching %r" % want)
print("warp instructions %d, of which DADD+DMUL+DFMA %d (%.1f%%), stall samples %d" % (
    total[0], total[1], 100.0 * total[1] / max(total[0], 1), total[2]))
print("opcode mix: " + ", ".join("%s %.1f%%" % (o, 100.0 * n / total[0])
                                 for o, n in sorted(ops.items(), key=lambda kv: -kv[1])[:24]))
by_file = collections.defaultdict(lambda: [0, 0, 0])
for (f, ln, text), v in per_line.items():
    for i in range(3):
        by_file[f][i] += v[i]
print("per file (inst% / fp64% of all inst / samples%):")
for f, v in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print("  %5.1f%% %5.1f%% %5.1f%%  %s" % (100.0 * v[0] / total[0], 100.0 * v[1] / total[0],
                                            100.0 * v[2] / max(total[2], 1), f))
print("source lines (inst% / fp64-inst% of all inst / samples%):")
for (f, ln, text), v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:topn]:
    print("  %5.2f%% %5.2f%% %5.2f%%  %s:%s  %s" % (100.0 * v[0] / total[0], 100.0 * v[1] / total[0],
                                                    100.0 * v[2] / max(total[2], 1), f, ln, text))
