#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): the slab-decomposed evaluation with
the NCCL ghost-plane exchange equals the single-GPU evaluation of the whole periodic domain,
bit for bit, for every workload family; fd_flag 0 and 1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from ampe_b200 import configs, fields, rhs
from ampe_b200.halo import DistributedRHS, slab_dim

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
SIZES = {"dendrite2d": dict(nx=96, ny=64 * world), "auni2d": dict(nx=96, ny=32 * world),
         "gg3d_hbsm": dict(nx=40, ny=24, nz=8 * world), "auni3d": dict(nx=40, ny=24, nz=8 * world),
         "pfhub1a": dict(nx=64, ny=16 * world)}
bad = 0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
_plog = open(os.path.join(ROOT, "gpurun_out", "mgpu_progress_rank%d.log" % rank), "a")


def progress(msg):
    _plog.write(msg + "\n")
    _plog.flush()


progress("start graphs=%s" % os.environ.get("AMPE_B200_GRAPHS"))
for name, kw in SIZES.items():
    cfg = configs.BUILDERS[name](**kw)
    cfg.symmetry_aware = 0
    st = fields.make_state(name, cfg)          # the whole domain, same on every rank
    yfull = rhs.to_device(st)
    rf = rhs.QuatIntegratorRHS(cfg, dev)
    kks = cfg.conc_rhs_form in (2, 3)
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
    c2.symmetry_aware = 0
    c2.nranks, c2.rank = world, rank
    dim = slab_dim(ndim)
    sl = slice(rank * ns, (rank + 1) * ns)
    cut = lambda t: (t[..., sl, :, :] if ndim == 3 else t[..., sl, :]).contiguous()
    y = rhs.SolutionVector({k: (None if v is None else cut(v)) for k, v in yfull.items()})
    r = rhs.QuatIntegratorRHS(c2, dev)
    drv = DistributedRHS(r, rank, world)
    if kks:
        c0 = y["conc"].reshape(-1).clone()
        drv.resetRefPhaseConcentrations(c0, c0.clone())
    for fd in (0, 1):
        out = y.like()
        # evaluations 1-2 run eagerly, the 3rd is captured into a CUDA graph, the 4th replays it
        for rep in range(4):
            if rep == 3:
                for v in out.values():
                    if v is not None:
                        v.fill_(float("nan"))
            progress("%s fd=%d rep=%d issue" % (name, fd, rep))
            drv.evaluateRHSFunction(0.0, y, out, fd)
            torch.cuda.synchronize()
            progress("%s fd=%d rep=%d done" % (name, fd, rep))
        for k, v in out.items():
            if v is None or (k == "quat" and not cfg.evolve_quat):
                continue
            same = torch.equal(v, cut(ref[fd][k]))
            if not same:
                bad += 1
                err = (v - cut(ref[fd][k])).abs().max().item()
                print("rank %d %s fd=%d %s MISMATCH max abs %.3e" % (rank, name, fd, k, err), flush=True)
    if rank == 0:
        print("%s: slab x%d == single GPU (graphs captured: %d, enabled: %s)" % (
            name, world, len(drv._graphs), drv.use_graphs), flush=True)
t = torch.tensor([bad], device=dev)
dist.all_reduce(t)
if rank == 0:
    print("MGPU CHECK", "OK" if t.item() == 0 else "FAILED (%d)" % t.item(), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if t.item() == 0 else 1)
