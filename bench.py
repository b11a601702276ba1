#!/usr/bin/env python
"""bench.py -- RHS cell-updates/s (GCUPS) of the fused evaluateRHSFunction path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one evaluateRHSFunction(fd_flag=0) over the whole (slab of the) grid.
Default workload = the configuration BASELINE.json's target is quoted on: examples/AuNi_3D
(phase + composition + quaternion RHS, CALPHAD KKS Newton per cell, EBS composition
flux) on a 1024x1024x128 slab per GPU, so that N = 8 is the 1024^3 grid weak-scaled over
8 B200 (slabs along z, ghost planes pushed into the neighbours' HBM over NVLink by
ampe_halo_*, overlapped with the interior evaluation).  Other workloads (--workload):
dendrite2d (2048^2), auni2d (4096^2, symmetry-aware), gg3d_hbsm (512x512x256 per GPU),
pfhub1a (200^2).

One JSON line is printed by rank 0 (contract in the task description):
  value     device-resident throughput (inputs already in HBM), whole job
  e2e       same metric through the host-buffer C-ABI call (H2D + D2H inside)
  roofline  algorithmic bytes of one evaluation / its CUDA-event time vs the measured HBM peak,
            with the per-kernel split (KKS pre-pass | fused kernel) measured by events on the
            launching stream (ampe_rhs_set_kernel_timing)
  cpu_baseline  the CPU restatement of AMPE's RHS (oracle, "port") on the host cores, on the
            sample config.cpu_sample names (same model, same synthetic recipe, smaller grid)
  newton    how the per-cell KKS Newton is started: the timed y is ONE TIME STEP AWAY from the
            state the reference concentrations were converged on (as in a run, where
            resetRefPhaseConcentrations is called once per step); the warm (same state) and cold
            (c_l = c_a = c) costs are reported beside it
--impl reference times the CPU restatement alone (the AMPE executable itself cannot be built
here: no gfortran/MPI/SAMRAI/SUNDIALS/Thermo4PFM -- see DESIGN.md), same model and recipe, on the
sample grid config.cpu_sample names, with all host threads."""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (builder kwargs per GPU, algorithmic bytes per cell per evaluation,
    #        (KKS pre-pass, fused kernel) split of those bytes: inputs read once, outputs written once per kernel)
    "dendrite2d": (dict(nx=2048, ny=2048), 72, (0, 72)),
    "auni2d": (dict(nx=4096, ny=4096), 136, (56, 120)),
    "gg3d_hbsm": (dict(nx=512, ny=512, nz=256), 96, (32, 112)),
    "auni3d": (dict(nx=1024, ny=1024, nz=128), 128, (56, 112)),
    "pfhub1a": (dict(nx=200, ny=200), 16, (0, 16)),
}
# grid of the CPU legs (cpu_baseline and --impl reference): same model and synthetic recipe; large enough
# to be far out of the host caches (>= 1 M cells x 20-60 arrays), small enough for seconds per evaluation
CPU_SAMPLE = {
    "dendrite2d": dict(nx=2048, ny=2048),
    "auni2d": dict(nx=1024, ny=1024),
    "gg3d_hbsm": dict(nx=256, ny=256, nz=64),
    "auni3d": dict(nx=256, ny=256, nz=128),
    "pfhub1a": dict(nx=200, ny=200),
}
MIN_TIMED_SECONDS = 0.3
MAX_PHASE_CHANGE = 0.01  # the timed state is y0 + dt ydot(y0) with dt such that max |d phi| is this


def dims(kw):
    return "x".join(str(v) for v in kw.values())


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


def base_config(args, world):
    kw, bytes_per_cell, _ = WORKLOADS[args.workload]
    ncell = 1
    for v in kw.values():
        ncell *= v
    newton = {"advanced": "Newton started one time step away from the converged reference state",
              "warm": "Newton started from the converged values of the same state",
              "cold": "cold Newton start (c_l = c_a = c)"}[args.newton]
    return {"workload": "%s %s per GPU, uniform periodic grid, fd_flag=%d, %s" % (
        args.workload, dims(kw), args.fd_flag, newton),
        "parallelism": "slab%d" % world if world > 1 else "single",
        "l2": "inputs larger than L2 (%.0f MB state per evaluation)" % (bytes_per_cell * 1e-6 * ncell),
        "cpu_sample": "%s %s (same model and synthetic recipe; grid of the CPU legs)" % (
            args.workload, dims(CPU_SAMPLE[args.workload]))}


def cpu_reference(workload, steps, warmup, newton, budget_s=25.0):
    """time the CPU restatement (oracle perf build, all host threads) on CPU_SAMPLE[workload]: the same model --
    symmetry-aware where the GPU arm is --, the same synthetic recipe and the same Newton start as the GPU arm"""
    import numpy as np
    from ampe_b200 import configs, fields
    from oracle import pyoracle
    kw = CPU_SAMPLE[workload]
    cfg = configs.BUILDERS[workload](**kw)
    st = fields.make_state(workload, cfg)
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg, perf=True)
    # all host threads this process may use (torchrun sets OMP_NUM_THREADS=1 for its workers)
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    o.L.oracle_set_num_threads(int(ncores))
    if cfg.symmetry_aware:
        n = o.ncell
        o.set_rotations([np.ones(n, dtype=np.int32) for _ in range(cfg.ndim)])
    kks = cfg.conc_rhs_form in (2, 3)
    ydot = o.alloc_like(y)
    if kks:
        o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
        if newton != "cold":
            o.eval(0.0, y, 0, ydot)
            o.set_ref(None, None)  # converged values of y0
    if newton == "advanced":
        o.eval(0.0, y, 0, ydot)
        if ydot.get("phase") is not None:
            dt = MAX_PHASE_CHANGE / max(float(np.abs(ydot["phase"]).max()), 1e-300)
        else:
            dt = MAX_PHASE_CHANGE / max(float(np.abs(ydot["conc"]).max()), 1e-300)
        for k in y:
            if y[k] is not None and ydot.get(k) is not None and (k != "quat" or cfg.evolve_quat):
                y[k] = y[k] + dt * ydot[k]
        if y.get("quat") is not None:
            q = y["quat"]
            y["quat"] = np.ascontiguousarray(q / np.sqrt((q * q).sum(0, keepdims=True)))
    for _ in range(max(1, warmup)):
        o.eval(0.0, y, 0, ydot)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        o.eval(0.0, y, 0, ydot)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt_wall = time.perf_counter() - t0
    ncell = o.ncell
    threads = o.L.oracle_num_threads()
    return {"value": ncell * done / dt_wall / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
            "sample": "%s %s (%d cells), %d evaluations (fd_flag=0, Newton: %s), oracle -O3 -march=native OpenMP, "
                      "%d threads" % (workload, dims(kw), ncell, done, newton, threads),
            "ms_per_step": dt_wall / done * 1e3, "steps": done}


def bind_near_gpu(prop):
    """restrict this process to the CPUs NVML reports as local to its GPU; returns how many, or None if the
    platform does not say (single NUMA node, container cpuset, no NVML)"""
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = "%08x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        h = nv.nvmlDeviceGetHandleByPciBusId(bus.encode())
        before = os.sched_getaffinity(0)
        nv.nvmlDeviceSetCpuAffinity(h)
        after = os.sched_getaffinity(0)
        if not after:
            os.sched_setaffinity(0, before)
            return None
        return len(after)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="auni3d", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ampe_b200", choices=["ampe_b200", "reference"])
    ap.add_argument("--fd-flag", type=int, default=0)
    ap.add_argument("--newton", default="advanced", choices=["advanced", "warm", "cold"],
                    help="start of the per-cell KKS Newton in the timed evaluations: 'advanced' (default) = the timed "
                         "state is one time step away from the state the reference concentrations were converged "
                         "on; 'warm' = same state (converges at the first residual check); 'cold' = c_l = c_a = c")
    ap.add_argument("--cold-ref", action="store_true", help="same as --newton cold")
    ap.add_argument("--no-lag", action="store_true",
                    help="A/B only: Integrator{lag_quat_sidegrad = FALSE} -- no lagged face arrays are kept, fd_flag=1 "
                         "evaluations recompute everything (the reference's default is TRUE)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the warm / cold Newton and per-kernel extras")
    args = ap.parse_args()
    if args.cold_ref:
        args.newton = "cold"
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    kw, bytes_per_cell, kernel_bytes = WORKLOADS[args.workload]
    config = base_config(args, world)

    if args.impl == "reference":
        if rank != 0:
            return
        ref = cpu_reference(args.workload, args.steps, args.warmup, args.newton, budget_s=120.0)
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
        # every rank's pinned host buffers and launching thread on the CPUs next to its GPU (first touch puts the
        # pages on that NUMA node): the end-to-end leg of N ranks shares the host's memory system.  Not at N = 1,
        # where the CPU baseline leg wants every host core.
        config["host_cpus_per_rank"] = bind_near_gpu(torch.cuda.get_device_properties(dev))
        dist.init_process_group("nccl", device_id=dev)

    cfg = configs.BUILDERS[args.workload](**kw)
    cfg.nranks, cfg.rank = world, rank
    if args.no_lag:
        cfg.lag_quat_sidegrad = 0
        config["workload"] += ", lag_quat_sidegrad off"
    st = fields.make_state(args.workload, cfg, device=dev, slab=(rank, world))
    y = rhs.SolutionVector(st)
    ydot = y.like()
    r = rhs.QuatIntegratorRHS(cfg, dev)
    drv = DistributedRHS(r, rank, world) if world > 1 else r
    if world > 1:
        config["halo_transport"] = drv.transport  # "ipc": ampe_halo_* (peer-mapped NVLink stores); "nccl": torch.distributed
        if drv.transport != "ipc" and cfg.symmetry_aware:
            # without the peer-mapped exchange the rotation indices' ghost planes cannot be fetched: say so
            cfg.symmetry_aware = 0
            config["workload"] += " (symmetry off: NCCL transport)"
            r = rhs.QuatIntegratorRHS(cfg, dev)
            drv = DistributedRHS(r, rank, world, transport="nccl")
    kks = cfg.conc_rhs_form in (2, 3)
    if cfg.symmetry_aware:
        n = r.ncell
        drv.setSymmetryRotations([torch.ones(n, dtype=torch.int32, device=dev) for _ in range(cfg.ndim)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def global_max(v):
        if world > 1:
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def set_ref(mode):
        """reference (initial-guess) phase concentrations: cold = c, else the converged values of y"""
        if not kks:
            return
        c0 = y["conc"].reshape(-1).clone()
        drv.resetRefPhaseConcentrations(c0, c0.clone())
        if mode != "cold":
            drv.evaluateRHSFunction(0.0, y, ydot, 0)
            torch.cuda.synchronize()
            drv.resetRefPhaseConcentrations()  # QuatModel::resetRefPhaseConcentrations: c_l, c_a of this state

    def timed(yv, nsteps, fd):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(nsteps):
            drv.evaluateRHSFunction(0.0, yv, ydot, fd)
        e1.record()
        barrier()
        return global_max(e0.elapsed_time(e1)) / nsteps

    set_ref("cold" if args.newton == "cold" else "warm")
    y_timed = y
    if args.newton == "advanced":
        # one explicit step of the size a CVODE step typically covers: max |d phi| = MAX_PHASE_CHANGE
        drv.evaluateRHSFunction(0.0, y, ydot, 0)
        lead = "phase" if ydot.get("phase") is not None else "conc"
        dt = MAX_PHASE_CHANGE / max(global_max(float(ydot[lead].abs().max().item())), 1e-300)
        y_timed = y.like()
        for k in rhs.COMPONENTS:
            if y.get(k) is not None:
                y_timed[k].copy_(y[k])
        r.linearSum(1.0, y, dt, ydot, y_timed)
        if y.get("quat") is not None:
            r.normalizeQuat(y_timed)
        config["newton_time_step"] = dt
    if world > 1:
        # set-up, before the W warm-up steps: the first exchanges map the peers' buffers
        for _ in range(3):
            drv.evaluateRHSFunction(0.0, y_timed, ydot, 0)
        barrier()
    ws = timed(y_timed, args.warmup, 0)  # W warm-up steps; their time sizes the timed region
    if args.fd_flag:
        drv.evaluateRHSFunction(0.0, y_timed, ydot, 0)
    repeats = max(1, int(math.ceil(MIN_TIMED_SECONDS / max(1e-6, ws * 1e-3 * args.steps))))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_per_step = timed(y_timed, args.steps * repeats, args.fd_flag)
    clocks = sampler.stop() if rank == 0 else None
    launches = drv.lastLaunchCount() * args.steps * repeats
    ncell_total = r.ncell * world
    value = ncell_total / (ms_per_step * 1e-3) / 1e9
    nf = r.newtonFailures()
    config["timed_region"] = "%d steps x %d repeats = %d evaluations, %.3f s" % (
        args.steps, repeats, args.steps * repeats, ms_per_step * 1e-3 * args.steps * repeats)

    # ---- roofline: one evaluation on one GPU (no exchange), CUDA events on the launching stream; per-kernel
    # split from the events the library records around its own launches
    peaks, peak_kind = measured_peaks()
    nk = max(3, min(args.steps, 20))
    r.setKernelTiming(True)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kk, kf = [], []
    torch.cuda.synchronize()
    k0.record()
    for _ in range(nk):
        r.evaluateRHSFunction(0.0, y_timed, ydot, args.fd_flag)
        a, b = r.lastKernelMs()
        kk.append(a)
        kf.append(b)
    k1.record()
    torch.cuda.synchronize()
    r.setKernelTiming(False)
    kms = (sum(kk) + sum(kf)) / nk
    achieved = bytes_per_cell * r.ncell / (kms * 1e-3) / 1e9
    kernels = []
    for name, msl, bpc in (("kks_kernel (per-cell KKS Newton pre-pass)", kk, kernel_bytes[0]),
                           ("fused RHS kernel (rhs_march_kernel / rhs_tile_kernel / ch_kernel)", kf, kernel_bytes[1])):
        t = sum(msl) / nk
        if t > 0 and bpc > 0:
            g = bpc * r.ncell / (t * 1e-3) / 1e9
            kernels.append({"kernel": name, "ms": t, "algorithmic_bytes_per_cell": bpc, "achieved": g,
                            "frac": g / peaks["hbm_gbs"]})
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "frac_of_nominal_8tbs": achieved / 8000.0,
                "traffic": None, "peak_kind": peak_kind,
                "kernel": "one evaluateRHSFunction on one GPU: %d launch(es), device time = sum of the kernels' "
                          "CUDA-event durations" % r.lastLaunchCount(),
                "algorithmic_bytes_per_cell": bytes_per_cell, "ms": kms, "kernels": kernels,
                "binding_limit": "fp64 pipe / issue slots, not HBM (profiles/README.md): see roofline.ncu"}
    tr = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(tr):
        try:
            tj = json.load(open(tr))
            roofline["traffic"] = tj.get("dram_bytes_per_launch")
            roofline["ncu"] = tj.get("ncu")  # FP64 pipe / issue-slot utilisation of the same kernels (offline capture)
            # the binding roof, live: FP64 operations of one evaluation (counted by ncu for this workload's per-GPU
            # grid: thread-level DADD + DMUL + 2 DFMA; the count scales with the cells) over the live device time,
            # against 148 SMs x 64 FP64 lanes x 2 (FMA) x SM clock
            ks = (tj.get("ncu") or {}).get("kernels") or []
            flop = sum(k["fp64_tflops"] * 1e12 * k["ncu_duration_ms"] * 1e-3 for k in ks)
            if flop > 0 and tj.get("cells"):
                flop *= r.ncell / float(tj["cells"])
                sm_hz = float((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
                peak = 148 * 64 * 2 * sm_hz / 1e12
                roofline["fp64"] = {"achieved": flop / (kms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                                    "frac": flop / (kms * 1e-3) / 1e12 / peak,
                                    "note": "FMA peak; the library is built --fmad=false (bitwise translation "
                                            "invariance), so most FP64 issue slots carry one flop: see "
                                            "roofline.ncu.kernels[].fp64_pipe_pct for the pipe utilisation"}
        except Exception:
            pass

    # ---- Newton extras: the same evaluation warm (zero iterations) and cold (c_l = c_a = c)
    newton = None
    if kks and not args.no_extras and args.fd_flag == 0:
        newton = {"timed": args.newton, args.newton + "_ms_per_step": ms_per_step}
        if args.newton != "warm":
            set_ref("warm")
            newton["warm_ms_per_step"] = timed(y, max(3, args.steps // 4), 0)
        if args.newton != "cold":
            set_ref("cold")
            newton["cold_ms_per_step"] = timed(y, 3, 0)
        # back to the state of the timed region for the end-to-end leg
        set_ref("cold" if args.newton == "cold" else "warm")

    # ---- e2e: host buffers through the C-ABI plugin call (H2D + kernels + D2H every step); at N > 1 every rank
    # moves its slab through the same chunk pipeline and the ghost planes travel device to device
    e2e = None
    if not args.no_e2e:
        yh, ydh = {}, {}
        h2d = d2h = 0
        for k in rhs.COMPONENTS:
            t = y_timed.get(k)
            yh[k] = None if t is None else t.cpu().pin_memory()
            ydh[k] = None if t is None else torch.empty_like(yh[k]).pin_memory()
            if t is not None:
                h2d += t.numel() * 8
                if k != "quat" or cfg.evolve_quat:
                    d2h += t.numel() * 8
        if world > 1 and drv.transport != "ipc":
            def host_step():  # NCCL transport: plain copies around the device evaluation
                for k in rhs.COMPONENTS:
                    if yh[k] is not None:
                        y_timed[k].copy_(yh[k], non_blocking=True)
                drv.evaluateRHSFunction(0.0, y_timed, ydot, 0)
                for k in rhs.COMPONENTS:
                    if ydh[k] is not None and (k != "quat" or cfg.evolve_quat):
                        ydh[k].copy_(ydot[k], non_blocking=True)
                torch.cuda.synchronize()
        else:
            def host_step():
                drv.evaluateRHSFunctionHost(0.0, yh, ydh, 0)
        for _ in range(2):
            host_step()
        n_e2e = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            host_step()
        barrier()
        dt_e = global_max((time.perf_counter() - t0) / n_e2e)
        e2e = {"value": r.ncell * world / dt_e / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": dt_e * 1e3, "steps": n_e2e,
               "path": "ampe_rhs_eval_host per rank: pinned host y -> device, evaluate, ydot -> pinned host, slab "
                       "chunks pipelined on three streams (H2D | kernels | D2H)" + (
                           "; ghost planes device to device (ampe_halo_*); max over ranks" if world > 1 else "")}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref = cpu_reference(args.workload, 1000, 1, args.newton, budget_s=15.0)
            cpu = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the checker is optional for the measurement itself
            cpu = {"value": None, "unit": "GCUPS", "cores": 0, "kind": "port", "sample": "failed: %r" % e}

    if rank == 0:
        line = {"metric": "RHS cell-updates/s", "value": value, "unit": "GCUPS", "n_gpus": world,
                "steps": args.steps, "repeats": repeats, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "newton": newton, "newton_failures": nf}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()  # nobody pushes into a neighbour's buffers any more
        drv.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
