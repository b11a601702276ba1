"""Parity at the sizes BASELINE.json quotes ("parity to green" (a), (c)): the fused path against the
restatement on the full Dendrite2D 2048^2 and AuNi_2D 4096^2 (symmetry-aware, random rotation indices) grids and
on 3D grids deep enough that every marching block (32 planes, ring of four staged planes) wraps its ring eight
times and stacks three blocks along z.  The checker is the parity build of the restatement with its cell loops
spread over the host threads (liboracle_par.so: same -O2 / no-contraction arithmetic per cell); the ill-conditioned
outputs are arbitrated by the long-double build (parity.check)."""
import os

import numpy as np
import pytest
import torch

import parity

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

CASES = {
    "dendrite2d": dict(nx=2048, ny=2048),            # C2 in full
    "auni2d": dict(nx=4096, ny=4096),                # C3 in full, Symmetry{} on
    "gg3d_hbsm": dict(nx=256, ny=128, nz=96),        # C4: 3 marching blocks along z, ring wrapped 8 x per block
    "auni3d": dict(nx=256, ny=128, nz=96),           # C5
}


@pytest.mark.parametrize("name", list(CASES))
def test_baseline_size_matches_oracle(name):
    from ampe_b200 import rhs
    from oracle import pyoracle
    cfg, st = parity.make_case(name, **CASES[name])
    rot = parity.random_rotations(cfg) if cfg.symmetry_aware else None
    y_np = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    fds = (0, 1)
    kks = cfg.conc_rhs_form in (2, 3)

    def run(o):
        if kks:
            o.set_ref(y_np["conc"].ravel().copy(), y_np["conc"].ravel().copy())
        if rot is not None:
            o.set_rotations(rot)
        res = []
        for fd in fds:
            status, yd = o.eval(0.0, y_np, fd)
            assert status == 0
            res.append(yd)
        extra = o.phase_concentrations() if kks else None
        o.close()
        return res, extra

    o_outs, o_extra = run(pyoracle.Oracle(cfg, perf="par"))
    ld = run(pyoracle.OracleLD(cfg)) if parity.needs_arbiter(cfg) else None
    g_outs, g_extra, launches = parity.run_gpu(cfg, st, fds, rot)
    errs = parity.compare_outputs(cfg, fds, o_outs, o_extra, g_outs, g_extra, ld)
    print(name, dict(errs), errs.ld)
    parity.check(errs)
