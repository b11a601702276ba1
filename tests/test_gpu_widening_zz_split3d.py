"""Experiment AMPE_B200_SPLIT3D=1 (rhs_march.cuh, PART 1 / 2): the 3D EBS evaluation as two lighter launches
-- phase + quaternion RHS, then composition RHS -- must reproduce the single fused launch BIT FOR BIT
(every output is computed by the same expressions), including the lagged face data that a following
fd_flag = 1 evaluation reads, on one slab and on split interior / boundary ranges.
Added at the end of round 1 without a GPU left: first executed by the round-end run."""
import os

import numpy as np
import pytest
import torch

import parity

# opt-in feature, opt-in tests: they join the default GPU suite once a GPU run has shown them green
# (tools/gpu_split3d.sh sets the variable)
# first executed on a B200 in round 2 (profiles/r02a_pytest_experiments.log): green, part of the default GPU suite
pytestmark = pytest.mark.gpu


def _run(split, cfg, st, fds, part_sequence=(0,)):
    from ampe_b200 import rhs
    old = os.environ.pop("AMPE_B200_SPLIT3D", None)
    if split:
        os.environ["AMPE_B200_SPLIT3D"] = "1"
    try:
        r = rhs.QuatIntegratorRHS(cfg)  # the switch is read when the context is created
    finally:
        os.environ.pop("AMPE_B200_SPLIT3D", None)
        if old is not None:
            os.environ["AMPE_B200_SPLIT3D"] = old
    y = rhs.to_device(st)
    c0 = y["conc"].reshape(-1).clone()
    r.resetRefPhaseConcentrations(c0, c0.clone())
    outs, launches = [], []
    for fd in fds:
        yd = y.like()
        for part in part_sequence:
            r.evaluateRHSFunction(0.0, y, yd, fd, part=part)
        torch.cuda.synchronize()
        launches.append(r.lastLaunchCount())
        outs.append({k: (None if v is None else v.cpu().numpy()) for k, v in yd.items()})
    assert r.newtonFailures() == 0
    r.close()
    return outs, launches


@pytest.mark.parametrize("kw", [None, dict(nx=40, ny=24, nz=35)])
def test_split_launches_are_bit_identical(kw):
    cfg, st = parity.make_case("auni3d", **(kw or {}))
    fds = (0, 1, 0)
    ref, l_ref = _run(False, cfg, st, fds)
    got, l_got = _run(True, cfg, st, fds)
    assert l_got[0] == l_ref[0] + 1  # KKS + two march launches instead of KKS + one
    for a, b in zip(ref, got):
        for k in ("phase", "quat", "conc"):
            assert np.array_equal(a[k], b[k]), k


def test_split_launches_with_interior_boundary_ranges():
    cfg, st = parity.make_case("auni3d")
    ref, _ = _run(False, cfg, st, (0, 1))
    got, _ = _run(True, cfg, st, (0, 1), part_sequence=(1, 2))
    for a, b in zip(ref, got):
        for k in ("phase", "quat", "conc"):
            assert np.array_equal(a[k], b[k]), k


def test_split_matches_oracle():
    cfg, st = parity.make_case("auni3d")
    os.environ["AMPE_B200_SPLIT3D"] = "1"
    try:
        errs = parity.compare("auni3d", cfg, st, fd_flags=(0, 1))
    finally:
        os.environ.pop("AMPE_B200_SPLIT3D", None)
    parity.check(errs)
