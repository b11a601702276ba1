#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU; also valid with several ranks on ONE GPU,
LOCAL_RANK all 0 -- CUDA IPC works between processes on the same device): the slab-decomposed evaluation
through the C-ABI exchange (ampe_halo_* / ampe_rhs_eval_slab) equals the single-GPU evaluation of the whole
periodic domain, bit for bit, for every workload family; fd_flag 0 and 1; the symmetry-aware path with
rotation indices whose ghost planes come from the neighbours; the host-buffer path (chunk pipeline + exchange);
with and without the overlap of exchange and interior evaluation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from ampe_b200 import configs, fields, rhs
from ampe_b200.halo import DistributedRHS, slab_dim

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ndev = torch.cuda.device_count()
local = int(os.environ.get("LOCAL_RANK", "0")) % ndev
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
backend = "nccl" if ndev >= world else "gloo"   # several ranks on one GPU: NCCL refuses duplicate devices
dist.init_process_group(backend, **({"device_id": dev} if backend == "nccl" else {}))
SIZES = {"dendrite2d": dict(nx=96, ny=64 * world), "auni2d": dict(nx=96, ny=32 * world),
         "gg3d_hbsm": dict(nx=40, ny=24, nz=8 * world), "auni3d": dict(nx=40, ny=24, nz=8 * world),
         "pfhub1a": dict(nx=64, ny=16 * world),
         # zero-slope physical boundaries in every direction: the exchange ring is cut at the first / last rank
         "dendrite2d:slope": dict(nx=96, ny=64 * world), "auni3d:slope": dict(nx=40, ny=24, nz=8 * world),
         "gg3d_hbsm:slope": dict(nx=40, ny=24, nz=8 * world)}
bad = 0


def rotations(cfg, seed=7):
    import parity
    return parity.random_rotations(cfg, seed)


MODES = (None,) if os.environ.get("MGPU_MODES") == "default" else (None, "0", "1")
for overlap in MODES:   # None: the library's own choice (in-kernel wait / overlap by message size)
    if overlap is None:
        os.environ.pop("AMPE_B200_HALO_OVERLAP", None)
    else:
        os.environ["AMPE_B200_HALO_OVERLAP"] = overlap
    for key, kw in SIZES.items():
        name, _, bc = key.partition(":")
        cfg = configs.BUILDERS[name](**kw)
        if bc:
            cfg.zero_slope[0] = cfg.zero_slope[1] = cfg.zero_slope[2] = 1
        st = fields.make_state(name, cfg)          # the whole domain, same on every rank
        yfull = rhs.to_device(st)
        rf = rhs.QuatIntegratorRHS(cfg, dev)
        kks = cfg.conc_rhs_form in (2, 3)
        rot = rotations(cfg) if cfg.symmetry_aware else None
        if rot is not None:
            rf.setSymmetryRotations([torch.as_tensor(a).to(dev) for a in rot])
        if kks:
            c0 = yfull["conc"].reshape(-1).clone()
            rf.resetRefPhaseConcentrations(c0, c0.clone())
        ref = [yfull.like(), yfull.like()]
        rf.evaluateRHSFunction(0.0, yfull, ref[0], 0)
        rf.evaluateRHSFunction(0.0, yfull, ref[1], 1)
        # my slab
        ndim = cfg.ndim
        ns = cfg.n[ndim - 1] // world
        kw2 = dict(kw)
        kw2["nz" if ndim == 3 else "ny"] = ns
        c2 = configs.BUILDERS[name](**kw2)
        for d in range(3):
            c2.dx[d] = cfg.dx[d]
        c2.nranks, c2.rank = world, rank
        for d in range(3):
            c2.zero_slope[d] = cfg.zero_slope[d]
        sl = slice(rank * ns, (rank + 1) * ns)
        cut = lambda t: (t[..., sl, :, :] if ndim == 3 else t[..., sl, :]).contiguous()
        y = rhs.SolutionVector({k: (None if v is None else cut(v)) for k, v in yfull.items()})
        r = rhs.QuatIntegratorRHS(c2, dev)
        drv = DistributedRHS(r, rank, world, transport="ipc")
        assert drv.transport == "ipc", "peer-mapped exchange unavailable"
        if rot is not None:
            shape = tuple(reversed([cfg.n[d] for d in range(ndim)]))
            drv.setSymmetryRotations([cut(torch.as_tensor(a).reshape(shape)).reshape(-1).to(dev) for a in rot])
        if kks:
            c0 = y["conc"].reshape(-1).clone()
            drv.resetRefPhaseConcentrations(c0, c0.clone())
        for fd in (0, 1):
            out = y.like()
            for rep in range(3):   # repeated: both buffer parities, flags running ahead
                if rep == 2:
                    for v in out.values():
                        if v is not None:
                            v.fill_(float("nan"))
                drv.evaluateRHSFunction(0.0, y, out, fd)
            torch.cuda.synchronize()
            for k, v in out.items():
                if v is None or (k == "quat" and not cfg.evolve_quat):
                    continue
                if not torch.equal(v, cut(ref[fd][k])):
                    bad += 1
                    err = (v - cut(ref[fd][k])).abs().max().item()
                    print("rank %d %s overlap=%s fd=%d %s MISMATCH max abs %.3e" % (rank, name, overlap, fd, k, err),
                          flush=True)
        if rot is not None:
            # the device pre-pass on the slab (ghost planes of y and of the indices from the neighbours) against
            # the pre-pass on the whole domain, then one evaluation with the indices it found
            rf.computeSymmetryRotations(yfull)
            drv.computeSymmetryRotations(y)
            shape = tuple(reversed([cfg.n[d] for d in range(ndim)]))
            for a, (tf, ts) in enumerate(zip(rf.symmetryRotations(), r.symmetryRotations())):
                if not torch.equal(ts, cut(tf.reshape(shape)).reshape(-1)):
                    bad += 1
                    print("rank %d %s rotation indices of direction %d MISMATCH" % (rank, name, a), flush=True)
            o1, o2 = yfull.like(), y.like()
            rf.evaluateRHSFunction(0.0, yfull, o1, 0)
            drv.evaluateRHSFunction(0.0, y, o2, 0)
            for k in ("phase", "quat", "conc"):
                if not torch.equal(o2[k], cut(o1[k])):
                    bad += 1
                    print("rank %d %s RHS with the pre-pass indices: %s MISMATCH" % (rank, name, k), flush=True)
            rf.evaluateRHSFunction(0.0, yfull, ref[0], 0)  # the host-path check below uses these indices too
        if overlap is None or overlap == "1":
            # time stepping: y changes at every step, so a stale ghost plane or a buffer reused too early shows up as
            # a difference from the single-rank trajectory (explicit Euler and Heun, Newton reference reset per step)
            import parity
            dt = 0.25 * parity.TRAJ_DT[name]   # (the grids here are finer than the ones the step sizes were found on)
            for scheme, nsteps in ((0, 12), (1, 5)):
                yf2 = rhs.SolutionVector({k: (None if v is None else v.clone()) for k, v in yfull.items()})
                ys2 = rhs.SolutionVector({k: (None if v is None else v.clone()) for k, v in y.items()})
                if kks:
                    c0f, c0s = yf2["conc"].reshape(-1).clone(), ys2["conc"].reshape(-1).clone()
                    rf.resetRefPhaseConcentrations(c0f, c0f.clone())
                    drv.resetRefPhaseConcentrations(c0s, c0s.clone())
                rf.integrateFixed(yf2, dt, nsteps, scheme=scheme)
                drv.integrateFixed(ys2, dt, nsteps, scheme=scheme)
                for k, v in ys2.items():
                    if v is None:
                        continue
                    if not torch.equal(v, cut(yf2[k])):
                        bad += 1
                        print("rank %d %s overlap=%s trajectory (scheme %d) %s MISMATCH max abs %.3e" % (
                            rank, name, overlap, scheme, k, (v - cut(yf2[k])).abs().max().item()), flush=True)
            if kks:   # back to the reference state of the checks below
                c0f, c0s = yfull["conc"].reshape(-1).clone(), y["conc"].reshape(-1).clone()
                rf.resetRefPhaseConcentrations(c0f, c0f.clone())
                drv.resetRefPhaseConcentrations(c0s, c0s.clone())
                rf.evaluateRHSFunction(0.0, yfull, ref[0], 0)
        # the host-buffer path of the same slab (fd_flag = 0)
        yh = {k: (None if v is None else v.cpu().pin_memory()) for k, v in y.items()}
        oh = {k: (None if v is None else torch.full_like(v.cpu(), float("nan")).pin_memory()) for k, v in y.items()}
        for chunks in ("1", "3"):
            os.environ["AMPE_B200_HOST_CHUNKS"] = chunks
            drv.evaluateRHSFunctionHost(0.0, yh, oh, 0)
            for k, v in oh.items():
                if v is None or (k == "quat" and not cfg.evolve_quat):
                    continue
                if not torch.equal(v, cut(ref[0][k]).cpu()):
                    bad += 1
                    print("rank %d %s host path (chunks %s) %s MISMATCH" % (rank, name, chunks, k), flush=True)
        os.environ.pop("AMPE_B200_HOST_CHUNKS", None)
        if rank == 0:
            print("%s: slab x%d == single GPU (overlap=%s, launches per evaluation %d)" % (
                key, world, overlap, drv.lastLaunchCount()), flush=True)
        drv.close()
        r.close()
        rf.close()
t = torch.tensor([bad], device=dev if backend == "nccl" else "cpu")
dist.all_reduce(t)
if rank == 0:
    print("MGPU CHECK", "OK" if t.item() == 0 else "FAILED (%d)" % t.item(), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if t.item() == 0 else 1)
