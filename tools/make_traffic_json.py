#!/usr/bin/env python
"""profiles/traffic_<workload>.json from an `ncu --set full` capture of one evaluation (the kernels of one
evaluateRHSFunction, captured by tools/gpu_r02i.sh): DRAM bytes per launch and per evaluation, FP64-pipe and issue
utilisation, achieved FP64 FLOP/s per kernel (thread-level DADD + DMUL + 2 DFMA).  bench.py copies
`dram_bytes_per_launch` into roofline.traffic and the `ncu` block into roofline.ncu.
usage: python tools/make_traffic_json.py <workload> <report.ncu-rep> <round tag>"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    workload, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}

    def val(r, name, scaled=False):
        v = float(r[col[name]].replace(",", ""))
        return v * scale[units[col[name]]] if scaled else v

    kernels, total = [], 0.0
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        rd, wr = val(r, "dram__bytes_read.sum", True), val(r, "dram__bytes_write.sum", True)
        t = val(r, "gpu__time_duration.sum", True)
        per_cycle = (val(r, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") +
                     val(r, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") +
                     2.0 * val(r, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"))
        cycles, ghz = val(r, "l1tex__cycles_elapsed.avg"), val(r, "l1tex__cycles_elapsed.avg.per_second")
        flops = per_cycle * cycles / t
        kernels.append({"kernel": name.split("(")[0][:90], "ncu_duration_ms": t * 1e3, "dram_read": rd, "dram_write": wr,
                        "fp64_pipe_pct": val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        "registers": val(r, "launch__registers_per_thread"),
                        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                        "fp64_tflops": flops / 1e12,
                        "fp64_tflops_frac_of_fma_peak": flops / (148 * 64 * 2 * ghz * 1e9)})
        total += rd + wr
    # one evaluation = the distinct kernels of the capture (a capture may hold the same kernel twice)
    seen, per_eval = set(), 0.0
    uniq = []
    for k in kernels:
        if k["kernel"] in seen:
            continue
        seen.add(k["kernel"])
        uniq.append(k)
        per_eval += k["dram_read"] + k["dram_write"]
    # cells of the capture = bench.py's default per-GPU grid of the workload (bench.py scales the FLOP count with it)
    cells = {"auni3d": 1024 * 1024 * 128, "dendrite2d": 2048 * 2048, "auni2d": 4096 * 4096,
             "gg3d_hbsm": 512 * 512 * 256, "pfhub1a": 200 * 200}.get(workload)
    d = {"workload": workload, "cells": cells, "dram_bytes_per_launch": per_eval,
         "kernel": " + ".join(k["kernel"].split("<")[0].replace("void ", "") for k in uniq) + " (one evaluateRHSFunction)",
         "source": "profiles/%s_ncu_full_%s.txt (ncu --set full --clock-control none, one launch per kernel)" % (tag, workload),
         "ncu": {"kernels": uniq, "fp64_fma_peak_tflops": 37.2,
                 "note": "FP64 FLOP/s = thread-level DADD + DMUL + 2 DFMA over the kernel's duration (tools/make_traffic_json.py)"}}
    path = os.path.join(ROOT, "profiles", "traffic_%s.json" % workload)
    json.dump(d, open(path, "w"), indent=1)
    print(path, "%.2f GB per evaluation" % (per_eval / 1e9), [(k["kernel"][:24], round(k["fp64_tflops"], 2)) for k in uniq])


if __name__ == "__main__":
    main()
