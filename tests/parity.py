"""Shared helpers of the parity tests: run the CUDA path (through the C ABI) and
the CPU oracle on the same seeded inputs and compare.

Error metric (SURVEY.md section 7 "Hard parts"): divergence-of-flux
cancellation makes a plain relative error meaningless where the RHS crosses
zero, so   err = |gpu - ref| / max(|ref|, FLOOR * ||ref||_inf)   with
FLOOR = 1e-3; the bar is the north_star's 1e-12 (fp64)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ampe_b200 import configs, fields  # noqa: E402

TOL = 1.0e-12
FLOOR = 1.0e-3

SMALL = {
    "pfhub1a": dict(nx=40, ny=36),
    "dendrite2d": dict(nx=72, ny=56),
    "auni2d": dict(nx=72, ny=56),
    "gg3d_hbsm": dict(nx=36, ny=20, nz=12),
    "auni3d": dict(nx=36, ny=20, nz=12),
}


def rel_err(gpu, ref):
    gpu = np.asarray(gpu, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    scale = np.maximum(np.abs(ref), FLOOR * max(np.abs(ref).max(), 1e-300))
    return float((np.abs(gpu - ref) / scale).max())


def random_rotations(cfg, seed=7):
    """synthetic quat_symm_rotation indices: mostly identity (1), some random
    +-(1..48) so that every rotation branch is exercised"""
    rng = np.random.default_rng(seed)
    n = cfg.n[0] * cfg.n[1] * (cfg.n[2] if cfg.ndim == 3 else 1)
    out = []
    for d in range(cfg.ndim):
        iq = np.ones(n, dtype=np.int32)
        pick = rng.random(n) < 0.3
        vals = rng.integers(1, 49, size=n) * rng.choice([-1, 1], size=n)
        iq[pick] = vals[pick]
        out.append(iq)
    return out


def make_case(name, **kw):
    cfg = configs.BUILDERS[name](**(kw or SMALL[name]))
    st = fields.make_state(name, cfg)
    if name in ("auni2d", "auni3d", "gg3d_hbsm"):
        # small grids: make sure several grains with different orientations exist
        phi, q = fields.grains(cfg, 5, max(3.0, cfg.n[0] / 9.0), 2.0)
        q = fields.smooth_unit(q)
        st["phase"], st["quat"] = phi, q
        h = fields.h_pbg(phi)
        c_in, c_out = (0.1, 0.06) if name == "gg3d_hbsm" else (0.096, 0.25)
        st["conc"] = (c_in * h + c_out * (1 - h) +
                      fields.smooth_noise(tuple(phi.shape), 1e-3, "cpu", 11)).contiguous()
    return cfg, st


def run_oracle(cfg, st, fd_flags=(0,), rotations=None, ref=None):
    from oracle import pyoracle
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    if cfg.conc_rhs_form in (2, 3):
        r = ref if ref is not None else (y["conc"].ravel().copy(), y["conc"].ravel().copy())
        o.set_ref(np.ascontiguousarray(r[0]), np.ascontiguousarray(r[1]))
    if cfg.symmetry_aware:
        o.set_rotations(rotations)
    outs = []
    for fd in fd_flags:
        status, yd = o.eval(0.0, y, fd_flag=fd)
        outs.append((status, yd))
    extra = o.phase_concentrations() if cfg.conc_rhs_form in (2, 3) else None
    o.close()
    return outs, extra


def run_gpu(cfg, st, fd_flags=(0,), rotations=None, ref=None, perturb=None):
    from ampe_b200 import rhs
    y = rhs.to_device(st)
    r = rhs.QuatIntegratorRHS(cfg)
    if cfg.conc_rhs_form in (2, 3):
        if ref is None:
            c0 = y["conc"].reshape(-1).clone()
            r.resetRefPhaseConcentrations(c0, c0.clone())
        else:
            r.resetRefPhaseConcentrations(torch.as_tensor(ref[0]).cuda(), torch.as_tensor(ref[1]).cuda())
    if cfg.symmetry_aware:
        r.setSymmetryRotations([torch.as_tensor(a).cuda() for a in rotations])
    outs = []
    for fd in fd_flags:
        yd = y.like()
        r.evaluateRHSFunction(0.0, y, yd, fd)
        torch.cuda.synchronize()
        outs.append({k: (None if v is None else v.cpu().numpy()) for k, v in yd.items()})
    extra = None
    if cfg.conc_rhs_form in (2, 3):
        cl, ca = r.phaseConcentrations()
        extra = (cl.cpu().numpy(), ca.cpu().numpy())
        nf = r.newtonFailures()
        assert nf == 0, "Newton failures on the GPU: %d" % nf
    launches = r.lastLaunchCount()
    r.close()
    return outs, extra, launches


def compare(name, cfg, st, fd_flags=(0,), rotations=None):
    o_outs, o_extra = run_oracle(cfg, st, fd_flags, rotations)
    g_outs, g_extra, _ = run_gpu(cfg, st, fd_flags, rotations)
    errs = {}
    for n, ((status, yo), yg) in enumerate(zip(o_outs, g_outs)):
        assert status == 0, "oracle Newton failure"
        for k in ("phase", "quat", "conc", "temperature"):
            if yo.get(k) is None:
                continue
            if k == "quat" and not cfg.evolve_quat:
                continue
            errs["fd%d:%s" % (fd_flags[n], k)] = rel_err(yg[k], yo[k])
    if o_extra is not None:
        errs["cl"] = rel_err(g_extra[0], o_extra[0])
        errs["ca"] = rel_err(g_extra[1], o_extra[1])
    return errs


# ---- fixed-step trajectories (north_star: field trajectories within 1e-8 after 100 steps) ----
# Explicit Euler step sizes per small case: about 1/5 of the empirical stability limit of the
# stiffest term on these grids (found with the oracle, tools/find_stable_dt.py).
TRAJ_DT = {
    "pfhub1a": 5.0e-3,     # limit ~2e-2
    "dendrite2d": 1.0e-9,  # limit ~5e-9 (singular orientation diffusivity 1/|grad q|)
    "auni2d": 5.0e-11,     # limit ~2e-10
    "gg3d_hbsm": 2.0e-8,   # limit ~1e-7
    "auni3d": 1.0e-10,     # limit ~5e-10
}


def normalize_quat_np(q):
    """QuatModel::normalizeQuat (QuatModel.cc:4237-4262), same operation order"""
    n2 = np.zeros_like(q[0])
    for m in range(q.shape[0]):
        n2 = n2 + q[m] * q[m]
    inv = 1.0 / np.sqrt(n2)
    return q * inv[None]


def oracle_trajectory(cfg, st, dt, nsteps, rotations=None, energy_every=0, scheme=0):
    """forward Euler (scheme 0) / Heun (1) with the CPU oracle; mirrors ampe_integrate_fixed"""
    from oracle import pyoracle
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    kks = cfg.conc_rhs_form in (2, 3)
    if kks:
        o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
    if cfg.symmetry_aware:
        o.set_rotations(rotations)
    evolved = [k for k in ("phase", "quat", "conc", "temperature")
               if y.get(k) is not None and not (k == "quat" and not cfg.evolve_quat)
               and not (k == "temperature" and not cfg.with_unsteady_temperature)]
    energies = []
    for s in range(nsteps):
        if energy_every and s % energy_every == 0:
            energies.append(o.energy(y)[1].copy())
        status, k1 = o.eval(0.0, y, 0)
        assert status == 0
        if scheme == 0:
            for k in evolved:
                y[k] = y[k] + dt * k1[k]
        else:
            ys = dict(y)
            for k in evolved:
                ys[k] = 1.0 * y[k] + dt * k1[k]
                y[k] = y[k] + (0.5 * dt) * k1[k]
            status, k2 = o.eval(0.0, ys, 0)
            assert status == 0
            for k in evolved:
                y[k] = y[k] + (0.5 * dt) * k2[k]
        if cfg.evolve_quat and cfg.qlen > 1:
            q = y["quat"]
            y["quat"] = np.ascontiguousarray(normalize_quat_np(q.reshape(cfg.qlen, -1)).reshape(q.shape))
        if kks and cfg.free_energy == 2:
            o.set_ref(None, None)
    if energy_every:
        energies.append(o.energy(y)[1].copy())
    o.close()
    return y, energies


def gpu_trajectory(cfg, st, dt, nsteps, rotations=None, energy_every=0, scheme=0):
    from ampe_b200 import rhs
    y = rhs.to_device(st)
    r = rhs.QuatIntegratorRHS(cfg)
    if cfg.conc_rhs_form in (2, 3):
        c0 = y["conc"].reshape(-1).clone()
        r.resetRefPhaseConcentrations(c0, c0.clone())
    if cfg.symmetry_aware:
        r.setSymmetryRotations([torch.as_tensor(a).cuda() for a in rotations])
    energies = []
    if energy_every:
        for s in range(0, nsteps, energy_every):
            energies.append(r.evaluateEnergy(y))
            r.integrateFixed(y, dt, min(energy_every, nsteps - s), scheme)
        energies.append(r.evaluateEnergy(y))
    else:
        r.integrateFixed(y, dt, nsteps, scheme)
    torch.cuda.synchronize()
    out = {k: (None if v is None else v.cpu().numpy()) for k, v in y.items()}
    if cfg.conc_rhs_form in (2, 3):
        assert r.newtonFailures() == 0
    r.close()
    return out, energies
