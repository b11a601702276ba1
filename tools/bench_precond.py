"""Timing of the block-preconditioner solve (csrc/mg.cu) on one GPU: ms per V-cycle and per level-0
half sweep at the BASELINE workload sizes, against the HBM roofline.  One JSON line per case.
Algorithmic bytes of the benchmarked block (M and C constants, D an array per direction --
the composition block): a sweep reads u, f and ND face arrays and writes u -> (3 + ND) * 8 B per cell (what the
fused tile pass moves; the two colour half-sweeps move twice that at sector granularity); a V(1,1) cycle = 2
sweeps + the fused residual-restriction ((2 + ND) * 8 B read per cell) + prolongation (16 B / cell)
on level 0, times
1 / (1 - 2^-ND) for the coarse levels.
usage (GPU box): python tools/bench_precond.py [--cases 2d:2048x2048,3d:256x256x256] [--cycles 10]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return json.load(open(p))["hbm_gbs"], "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback"


def main():
    from ampe_b200.precond import LevelSolver
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="2d:2048x2048,2d:4096x4096,3d:256x256x256,3d:512x512x512")
    ap.add_argument("--cycles", type=int, default=10)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--quat", action="store_true",
                    help="the quaternion block: column scale, four components in one solver (ampe_mg_create_multi)")
    ap.add_argument("--stream", action="store_true", help="launch on a non-default stream (needed by AMPE_B200_MG_GRAPH=1)")
    a = ap.parse_args()
    hbm, src = peak()
    for case in a.cases.split(","):
        kind, dims = case.split(":")
        n = [int(v) for v in dims.split("x")]
        nd = len(n)
        ncell = 1
        for v in n:
            ncell *= v
        shape = (n[2] if nd == 3 else 1, n[1], n[0])
        torch.manual_seed(1)
        ncomp = 4 if a.quat else 1
        g = LevelSolver(n, [1.0] * nd, with_column_scale=a.quat, ncomp=ncomp)
        sides = []
        for ax in range(nd):
            s = list(shape)
            s[2 - ax] += 1
            sides.append(-(50.0 + 10.0 * torch.rand(s, dtype=torch.float64, device="cuda")))
        if a.quat:
            gshape = tuple(v + 2 if (2 - ax) < nd else v for ax, v in enumerate(shape))
            mob = 0.5 + torch.rand(gshape, dtype=torch.float64, device="cuda")
            g.set_quat(0.01, mob, 1, sides, 0)
            rhs = torch.randn((ncomp,) + shape, dtype=torch.float64, device="cuda")
        else:
            g.set_elliptic(m_const=1.0, c_const=1.0, d=sides, ngd=0)
            rhs = torch.randn(shape, dtype=torch.float64, device="cuda")
        out = torch.empty_like(rhs)
        stream = torch.cuda.Stream() if a.stream else None
        g.solve(rhs, ncycles=2, out=out, symmetrized=a.quat)
        torch.cuda.synchronize()
        res0 = None if a.quat else float((rhs - g.apply(out)).norm() / rhs.norm())
        best = None
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # the library launches on the legacy default stream (stream = NULL), as torch does, unless --stream
            e0.record(stream)
            g.solve(rhs, ncycles=a.cycles, out=out, stream=stream, symmetrized=a.quat)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        launches = g.last_launch_count()
        per_cycle = best / a.cycles
        if a.quat:  # per cell: ncomp x (u in, u out, f) + the shared M, S and ND face arrays
            lvl0 = 2 * (ncomp * 24 + (2 + nd) * 8) + (ncomp * 16 + (2 + nd) * 8) + ncomp * 16
        else:
            lvl0 = 2 * (3 + nd) * 8 + (2 + nd) * 8 + 16
        bytes_cycle = ncell * lvl0 / (1.0 - 0.5 ** nd)
        gbs = bytes_cycle / (per_cycle * 1e-3) / 1e9
        print(json.dumps({"case": case, "block": "quaternion x4" if a.quat else "scalar", "levels": g.num_levels(), "ms_per_vcycle": per_cycle,
                          "launches_per_solve": launches, "algorithmic_bytes_per_vcycle": bytes_cycle,
                          "achieved_gbs": gbs, "hbm_peak_gbs": hbm, "peak_source": src, "frac": gbs / hbm,
                          "rel_residual_after_2_cycles": res0, "cells": ncell,
                          "tail": os.environ.get("AMPE_B200_MG_TAIL", "1"), "graph": os.environ.get("AMPE_B200_MG_GRAPH", "0"),
                          "fused": os.environ.get("AMPE_B200_MG_FUSED", "1")}))
        g.close()


if __name__ == "__main__":
    main()
