#!/usr/bin/env python
"""Quick GPU parity + timing sweep (development aid; the judged tests are tests/)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import parity
from ampe_b200 import configs, fields, rhs

for name in ("pfhub1a", "dendrite2d", "gg3d_hbsm", "auni3d", "auni2d"):
    cfg, st = parity.make_case(name)
    rot = None
    if cfg.symmetry_aware:
        rot = parity.random_rotations(cfg)
    try:
        errs = parity.compare(name, cfg, st, fd_flags=(0, 1, 0), rotations=rot)
        print(name, {k: "%.2e" % v for k, v in errs.items()}, flush=True)
    except Exception as e:
        print(name, "FAILED", repr(e), flush=True)

BIG = {"dendrite2d": dict(nx=2048, ny=2048), "auni2d": dict(nx=4096, ny=4096),
       "gg3d_hbsm": dict(nx=512, ny=512, nz=256), "auni3d": dict(nx=512, ny=512, nz=128)}
for name, kw in BIG.items():
    cfg = configs.BUILDERS[name](**kw)
    if name == "auni2d":
        cfg.symmetry_aware = 0
    st = fields.make_state(name, cfg, device="cuda")
    y = rhs.SolutionVector(st)
    r = rhs.QuatIntegratorRHS(cfg)
    if cfg.conc_rhs_form in (2, 3):
        c0 = y["conc"].reshape(-1).clone()
        r.resetRefPhaseConcentrations(c0, c0.clone())
    yd = y.like()
    for fd in (0, 1):
        for _ in range(3):
            r.evaluateRHSFunction(0.0, y, yd, fd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 10
        for _ in range(K):
            r.evaluateRHSFunction(0.0, y, yd, fd)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print("%s %s fd=%d: %.3f ms/eval, %.2f GCUPS, newton failures %d" % (
            name, kw, fd, ms, r.ncell / ms / 1e6, r.newtonFailures()), flush=True)
    if cfg.conc_rhs_form in (2, 3):
        # warm start: ref = converged values
        r.resetRefPhaseConcentrations()
        for _ in range(2):
            r.evaluateRHSFunction(0.0, y, yd, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            r.evaluateRHSFunction(0.0, y, yd, 0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print("%s warm-ref fd=0: %.3f ms/eval, %.2f GCUPS" % (name, ms, r.ncell / ms / 1e6), flush=True)
    r.close()
    del y, yd, st
    torch.cuda.empty_cache()
