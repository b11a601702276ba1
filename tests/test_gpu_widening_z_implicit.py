"""SURVEY.md 8f rank 1 on the device: the implicit integrator (ampe_b200/host/ImplicitIntegrator.h) driving
device vectors through the C ABI -- ampe_rhs_eval with fd_flag = 0 / 1, ampe_vec_*, ampe_apply_projection,
ampe_normalize_quat -- against the SAME integrator template driven by the CPU oracle.  Tight integrator
tolerances keep the two runs on the same iteration path; the bar is the north_star's 1e-8 on the fields."""
import os

import numpy as np
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu

TIGHT = dict(order=2, rtol=1e-8, atol=1e-10, max_krylov=30, max_newton=8)
# step = mult x the explicit-Euler step of the trajectory tests (1/5 of the stability limit)
CASES = [("pfhub1a", 40, 20), ("dendrite2d", 50, 10), ("auni2d", 20, 6), ("gg3d_hbsm", 10, 5), ("auni3d", 20, 4)]


def _oracle_run(name, dt, nsteps, rot, **kw):
    from oracle import pyoracle
    cfg, st = parity.make_case(name)
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    if cfg.conc_rhs_form in (2, 3):
        o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
    if rot is not None:
        o.set_rotations(rot)
    rc, stats = o.integrate_implicit(y, dt, nsteps, **kw)
    o.close()
    return y, rc, stats


def _device_run(name, dt, nsteps, rot, **kw):
    from ampe_b200 import rhs
    from ampe_b200.host_rhs import HostQuatIntegrator
    cfg, st = parity.make_case(name)
    y = rhs.to_device(st)
    h = HostQuatIntegrator(cfg, True)
    if cfg.conc_rhs_form in (2, 3):
        c0 = y["conc"].reshape(-1).clone()
        h.resetRefPhaseConcentrations(c0, c0.clone())
    if rot is not None:
        h.setSymmetryRotations([torch.as_tensor(a).cuda() for a in rot])
    rc, stats = h.integrateImplicit(y, dt, nsteps, **kw)
    torch.cuda.synchronize()
    out = {k: (None if v is None else v.cpu().numpy()) for k, v in y.items()}
    h.close()
    return cfg, st, out, rc, stats


@pytest.mark.parametrize("name,mult,nsteps", CASES)
def test_implicit_trajectory_matches_oracle_backend(name, mult, nsteps):
    dt = parity.TRAJ_DT[name] * mult
    cfg0, _ = parity.make_case(name)
    rot = parity.random_rotations(cfg0) if cfg0.symmetry_aware else None
    yo, rc_o, so = _oracle_run(name, dt, nsteps, rot, **TIGHT)
    cfg, st, yg, rc_g, sg = _device_run(name, dt, nsteps, rot, **TIGHT)
    assert rc_o == 0 and rc_g == 0, (so, sg)
    assert sg["steps"] == nsteps and sg["projections"] == nsteps
    assert sg["jtimes_evals"] == sg["linear_iterations"] > 0
    for k in ("phase", "quat", "conc", "temperature"):
        if yo.get(k) is None:
            continue
        scale = max(np.abs(yo[k]).max(), 1e-300)
        assert np.abs(yg[k] - yo[k]).max() <= 1e-8 * scale, (k, sg, so)
        moved = np.abs(yo[k] - st[k].numpy()).max()
        if moved > 1e-4 * scale:  # the increment itself agrees, not just the (large) field
            assert np.abs(yg[k] - yo[k]).max() <= 1e-3 * moved, (k, moved)
    if cfg.evolve_quat and cfg.qlen > 1:
        q = yg["quat"].reshape(cfg.qlen, -1)
        assert np.abs((q * q).sum(0) - 1.0).max() < 1e-14


def test_default_options_run():
    """AMPE's default integrator options (QuatIntegrator.cc:285-301) on the device"""
    dt = parity.TRAJ_DT["dendrite2d"]
    cfg, st, y, rc, stats = _device_run("dendrite2d", 20 * dt, 5, None)
    assert rc == 0 and stats["steps"] == 5


def test_newton_failure_code():
    """a step far beyond what two Newton iterations can absorb comes back as IMPLICIT_ENEWTON instead of a
    silent wrong answer (the CPU run of the same template returns the same code)"""
    from ampe_b200.host_rhs import HostQuatIntegrator
    dt = parity.TRAJ_DT["dendrite2d"]
    kw = dict(order=1, rtol=1e-10, atol=1e-12, max_krylov=3, max_newton=2)
    _, rc_o, _ = _oracle_run("dendrite2d", 1e4 * dt, 1, None, **kw)
    cfg, st, y, rc, stats = _device_run("dendrite2d", 1e4 * dt, 1, None, **kw)
    assert rc_o == HostQuatIntegrator.IMPLICIT_ENEWTON
    assert rc == HostQuatIntegrator.IMPLICIT_ENEWTON
