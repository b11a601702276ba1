#!/usr/bin/env python
"""Summarise an ncu report's source page (needs -lineinfo + --import-source on):
 (1) executed warp instructions per SASS opcode, (2) per source line of our files.
Usage: ncu_source_hot.py report.ncu-rep [top-N]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 45


def page(view):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


# ---- SASS opcodes (first kernel in the report only)
rows = page("sass")
hdr, nk = None, 0
op = collections.defaultdict(lambda: [0, 0])
tot = [0, 0]
for r in rows:
    if r and r[0] == "Kernel Name":
        nk += 1
        if nk > 1:
            break
        print(r[1][:150])
        continue
    if r and r[0] == "Address":
        hdr = r
        ii, si, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
        continue
    if hdr is None or len(r) <= max(ii, si):
        continue
    try:
        n, s = int(r[ii]), int(r[si])
    except ValueError:
        continue
    w = r[src].split()
    o = w[1] if (w and w[0].startswith("@") and len(w) > 1) else (w[0] if w else "?")
    o = o.split(".")[0]
    op[o][0] += n
    op[o][1] += s
    tot[0] += n
    tot[1] += s
print("SASS: total warp-inst %d, samples %d" % (tot[0], tot[1]))
for k, v in sorted(op.items(), key=lambda kv: -kv[1][0])[:25]:
    print("  %-10s inst %6.2f%%  samples %6.2f%%" % (k, 100.0 * v[0] / tot[0], 100.0 * v[1] / max(tot[1], 1)))

# ---- source lines
rows = page("cuda,sass")
cur, fn, seen_fn = None, None, set()
lines = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        si = hdr.index("# Samples")
        continue
    if hdr is None or r[0] == "" or len(r) <= max(ii, si):
        continue
    try:
        lines.append((cur, int(r[0]), r[1].strip(), int(r[ii]), int(r[si]), fn))
    except ValueError:
        pass
f0 = lines[0][5] if lines else None
lines = [l for l in lines if l[5] == f0]
ti = sum(l[3] for l in lines) or 1
ts = sum(l[4] for l in lines) or 1
print("SOURCE lines (inst%% / samples%%), attributed total inst %d" % ti)
for l in sorted(lines, key=lambda l: -l[3])[:topn]:
    print("  %5.2f%% %5.2f%%  %s:%d  %s" % (100.0 * l[3] / ti, 100.0 * l[4] / ts, l[0], l[1], l[2][:100]))
