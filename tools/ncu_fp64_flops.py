"""Achieved FP64 FLOP/s per kernel from an `ncu --set full` report (north_star: "achieved FP64 FLOP/s for the
Newton kernel"): thread-level DADD + DMUL + 2 DFMA (predicated-on, summed over the SMSPs, per elapsed cycle)
times the elapsed SM cycles, over the kernel's duration.  Peak for the fraction: 148 SMs x 64 FP64 lanes x 2
(FMA) x the SM clock of the capture.
usage: python tools/ncu_fp64_flops.py gpurun_out/prof_auni3d.ncu-rep [...]"""
import csv
import subprocess
import sys


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}

        def val(r, name):
            return float(r[col[name]].replace(",", ""))

        for r in rows[2:]:
            per_cycle = (val(r, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") +
                         val(r, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") +
                         2.0 * val(r, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"))
            cycles = val(r, "l1tex__cycles_elapsed.avg")  # SM-domain elapsed cycles
            t = val(r, "gpu__time_duration.sum") * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[units[col["gpu__time_duration.sum"]]]
            ghz = val(r, "l1tex__cycles_elapsed.avg.per_second")
            peak = 148 * 64 * 2 * ghz * 1e9
            flops = per_cycle * cycles / t
            print("%s | %s | %.3f ms | %.2f TFLOP/s fp64 | %.1f %% of %.1f TFLOP/s (148 SM x 64 lanes x 2 x %.3f GHz) | "
                  "FP64 pipe %.1f %% | %.0f flop/cycle" %
                  (rep.split("/")[-1], r[col["Kernel Name"]][:60], t * 1e3, flops / 1e12, 100 * flops / peak, peak / 1e12,
                   ghz, val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), per_cycle))


if __name__ == "__main__":
    main()
