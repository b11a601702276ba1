#!/usr/bin/env python
"""bench.py -- RHS cell-updates/s (GCUPS) of the fused evaluateRHSFunction path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one evaluateRHSFunction(fd_flag=0) over the whole (slab of the) grid.
Default workload = BASELINE.json configs[1]: examples/Dendrite2D (anisotropic phase
+ quaternion + temperature RHS) on a 2048x2048 uniform periodic grid per GPU; with
N>1 ranks the domain is N slabs along y (weak scaling, NCCL halo exchange
overlapped with the interior evaluation).  Other workloads: auni2d (4096^2),
gg3d_hbsm (512x512x256 per GPU), auni3d (1024x1024x128 per GPU), pfhub1a (200^2).

One JSON line is printed by rank 0 (contract in the task description):
  value     device-resident throughput (inputs already in HBM), whole job
  e2e       same metric through the host-buffer C-ABI call (H2D + D2H inside)
  roofline  dominant kernel: algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline  the CPU restatement of AMPE's RHS (oracle, "port") on the host cores
--impl reference times that CPU restatement alone (the AMPE executable itself cannot
be built here: no gfortran/MPI/SAMRAI/SUNDIALS/Thermo4PFM -- see DESIGN.md)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (builder kwargs per GPU, algorithmic bytes per cell per evaluation)
    "dendrite2d": (dict(nx=2048, ny=2048), 72),
    "auni2d": (dict(nx=4096, ny=4096), 136),
    "gg3d_hbsm": (dict(nx=512, ny=512, nz=256), 96),
    "auni3d": (dict(nx=1024, ny=1024, nz=128), 128),
    "pfhub1a": (dict(nx=200, ny=200), 16),
}
CPU_SAMPLE = {
    "dendrite2d": dict(nx=1024, ny=1024),
    "auni2d": dict(nx=512, ny=512),
    "gg3d_hbsm": dict(nx=96, ny=96, nz=64),
    "auni3d": dict(nx=96, ny=96, nz=64),
    "pfhub1a": dict(nx=200, ny=200),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (NVML polled every 2 ms in a
    thread; the nvidia-smi CLI is too slow for millisecond regions)."""

    def __init__(self, gpu_index):
        self.samples, self.reasons = [], set()
        self.gpu, self.run, self.t, self.h = gpu_index, False, None, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            self.h = None
            return
        self.run = True
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def _poll(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                "hw_thermal_slowdown": 0x40}
        while self.run:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = get_reasons(self.h)
                for n, b in bits.items():
                    if r & b:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        self.run = False
        self.t.join(timeout=1)
        sm = sorted(self.samples)
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def cpu_reference(workload, steps, warmup, budget_s=25.0):
    """time the CPU restatement (oracle perf build, all host threads) on a bounded sample"""
    import numpy as np
    from ampe_b200 import configs, fields
    from oracle import pyoracle
    kw = CPU_SAMPLE[workload]
    cfg = configs.BUILDERS[workload](**kw)
    if workload == "auni2d":
        cfg.symmetry_aware = 0
    st = fields.make_state(workload, cfg)
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg, perf=True)
    # all host threads this process may use (torchrun sets OMP_NUM_THREADS=1 for its workers)
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    o.L.oracle_set_num_threads(int(ncores))
    if cfg.conc_rhs_form in (2, 3):
        o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
        o.eval(0.0, y)
        o.set_ref(None, None)  # warm start from converged values
    ydot = o.alloc_like(y)
    for _ in range(max(1, warmup)):
        o.eval(0.0, y, 0, ydot)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        o.eval(0.0, y, 0, ydot)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    ncell = o.ncell
    threads = o.L.oracle_num_threads()
    return {"value": ncell * done / dt / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
            "sample": "%s %s, %d evaluations (fd_flag=0), oracle -O3 -march=native OpenMP" % (
                workload, "x".join(str(v) for v in kw.values()), done),
            "ms_per_step": dt / done * 1e3, "steps": done}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="dendrite2d", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ampe_b200", choices=["ampe_b200", "reference"])
    ap.add_argument("--fd-flag", type=int, default=0)
    ap.add_argument("--cold-ref", action="store_true",
                    help="KKS Newton starts from c_l_ref = c_a_ref = c in every evaluation (default: warm start "
                         "from the converged values, as after QuatModel::Advance)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    kw, bytes_per_cell = WORKLOADS[args.workload]
    config = {"workload": "%s %s per GPU, uniform periodic grid, fd_flag=%d%s" % (
        args.workload, "x".join(str(v) for v in kw.values()), args.fd_flag,
        ", cold Newton start" if args.cold_ref else ""),
        "parallelism": "slab%d" % world if world > 1 else "single",
        "l2": "inputs larger than L2 (%.0f MB state per evaluation)" % (
            bytes_per_cell * 1e-6 * eval("*".join(str(v) for v in kw.values())))}

    if args.impl == "reference":
        if rank != 0:
            return
        ref = cpu_reference(args.workload, args.steps, args.warmup, budget_s=120.0)
        line = {"metric": "RHS cell-updates/s", "value": ref["value"], "unit": "GCUPS",
                "n_gpus": args.gpus, "steps": ref["steps"], "warmup": args.warmup,
                "ms_per_step": ref["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "impl": "reference",
                "cpu_baseline": {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": ref["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from ampe_b200 import configs, fields, rhs
    from ampe_b200.halo import DistributedRHS

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL's internal stream at high priority: the ghost-plane exchange overlaps the interior
        # kernel instead of queueing behind its thread blocks
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = configs.BUILDERS[args.workload](**kw)
    cfg.nranks, cfg.rank = world, rank
    if args.workload == "auni2d" and world > 1:
        cfg.symmetry_aware = 0
    st = fields.make_state(args.workload, cfg, device=dev, slab=(rank, world))
    y = rhs.SolutionVector(st)
    ydot = y.like()
    r = rhs.QuatIntegratorRHS(cfg, dev)
    drv = DistributedRHS(r, rank, world) if world > 1 else r
    if cfg.symmetry_aware:
        n = r.ncell
        r.setSymmetryRotations([torch.ones(n, dtype=torch.int32, device=dev) for _ in range(cfg.ndim)])
    if cfg.conc_rhs_form in (2, 3):
        c0 = y["conc"].reshape(-1).clone()
        drv.resetRefPhaseConcentrations(c0, c0.clone())
        drv.evaluateRHSFunction(0.0, y, ydot, 0)
        torch.cuda.synchronize()
        if args.cold_ref:
            pass  # keep c_l_ref = c_a_ref = c: several Newton iterations per cell and evaluation
        elif world > 1:
            cl, ca = r.phaseConcentrations()
            drv.resetRefPhaseConcentrations(cl, ca)
        else:
            r.resetRefPhaseConcentrations()  # warm start, as after QuatModel::Advance

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:
        # set-up, before the W warm-up steps: NCCL builds its channels lazily during the first
        # exchanges (at N=2 a 20-step timing right after 3 evaluations read 0.31 ms/step, the same
        # code 0.17 ms/step over 100 steps: profiles/README.md)
        for _ in range(30):
            drv.evaluateRHSFunction(0.0, y, ydot, 0)
        barrier()
        config["setup_evaluations_before_warmup"] = 30
    for _ in range(args.warmup):
        drv.evaluateRHSFunction(0.0, y, ydot, 0)
    if args.fd_flag:
        drv.evaluateRHSFunction(0.0, y, ydot, args.fd_flag)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        drv.evaluateRHSFunction(0.0, y, ydot, args.fd_flag)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = r.lastLaunchCount() * args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    ncell_total = r.ncell * world
    value = ncell_total / (ms_per_step * 1e-3) / 1e9
    nf = r.newtonFailures()

    # ---- dominant kernel alone: the fused kernel is the only (or last) launch of an evaluation;
    # time it with CUDA events on the launching stream over the same number of launches
    peaks, peak_kind = measured_peaks()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(args.steps):
        r.evaluateRHSFunction(0.0, y, ydot, args.fd_flag)
    k1.record()
    torch.cuda.synchronize()
    kms = k0.elapsed_time(k1) / args.steps
    achieved = bytes_per_cell * r.ncell / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "frac_of_nominal_8tbs": achieved / 8000.0,
                "traffic": None, "peak_kind": peak_kind,
                "kernel": "one evaluateRHSFunction on one GPU (%d launch(es))" % r.lastLaunchCount(),
                "algorithmic_bytes_per_cell": bytes_per_cell}
    tr = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(tr):
        try:
            tj = json.load(open(tr))
            roofline["traffic"] = tj.get("dram_bytes_per_launch")
            roofline["ncu"] = tj.get("ncu")  # FP64 pipe / issue-slot utilisation of the same kernel (offline capture)
        except Exception:
            pass

    # ---- e2e: host buffers through the C-ABI plugin call (H2D + kernel + D2H every step)
    e2e = None
    if not args.no_e2e and world == 1:
        yh, ydh = {}, {}
        h2d = d2h = 0
        for k in rhs.COMPONENTS:
            t = y.get(k)
            yh[k] = None if t is None else t.cpu().pin_memory()
            ydh[k] = None if t is None else torch.empty_like(yh[k]).pin_memory()
            if t is not None:
                h2d += t.numel() * 8
                if k != "quat" or cfg.evolve_quat:
                    d2h += t.numel() * 8
        for _ in range(2):
            r.evaluateRHSFunctionHost(0.0, yh, ydh, 0)
        n_e2e = max(3, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            r.evaluateRHSFunctionHost(0.0, yh, ydh, 0)
        dt = (time.perf_counter() - t0) / n_e2e
        e2e = {"value": r.ncell / dt / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3, "steps": n_e2e,
               "path": "ampe_rhs_eval_host: pinned host y -> device, evaluate, ydot -> pinned host, slab chunks pipelined on three streams (H2D | kernels | D2H)"}
    elif not args.no_e2e and world > 1:
        # N ranks: every rank moves its slab host -> device, exchanges ghost planes with its
        # neighbours (NCCL, overlapped with the interior planes), evaluates, and reads ydot back
        yh, ydh = {}, {}
        h2d = d2h = 0
        for k in rhs.COMPONENTS:
            t = y.get(k)
            yh[k] = None if t is None else t.cpu().pin_memory()
            ydh[k] = None if t is None else torch.empty_like(yh[k]).pin_memory()
            if t is not None:
                h2d += t.numel() * 8
                if k != "quat" or cfg.evolve_quat:
                    d2h += t.numel() * 8

        def e2e_step():
            for k in rhs.COMPONENTS:
                if yh[k] is not None:
                    y[k].copy_(yh[k], non_blocking=True)
            drv.evaluateRHSFunction(0.0, y, ydot, 0)
            for k in rhs.COMPONENTS:
                if ydh[k] is not None and (k != "quat" or cfg.evolve_quat):
                    ydh[k].copy_(ydot[k], non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            e2e_step()
        n_e2e = max(3, min(args.steps, 20))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": r.ncell * world / dt / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": dt * 1e3, "steps": n_e2e,
               "path": "per rank: pinned host slab -> device, NCCL ghost-plane exchange overlapped with "
                       "the interior evaluation, ydot -> pinned host (max over ranks)"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref = cpu_reference(args.workload, 1000, 1, budget_s=15.0)
            cpu = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the checker is optional for the measurement itself
            cpu = {"value": None, "unit": "GCUPS", "cores": 0, "kind": "port", "sample": "failed: %r" % e}

    if rank == 0:
        line = {"metric": "RHS cell-updates/s", "value": value, "unit": "GCUPS", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "newton_failures": nf}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
