"""Pin the oracle's energy restatement (oracle/energy.cc: quatenergy / phi_interface_energy /
bulkenergy of {2d,3d}/quatenergy.m4, and the PFHub-1a functional) with closed-form cases and
discrete identities -- the reference ships no known-answer test for evaluateEnergy, so these are
properties of the formulas themselves.  CPU only."""
import numpy as np
import pytest

import parity
from oracle import pyoracle


def _np_state(st):
    return {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}


def _volume(cfg):
    v = 1.0
    for d in range(cfg.ndim):
        v *= cfg.n[d] * cfg.dx[d]
    return v


@pytest.mark.parametrize("name", ["dendrite2d", "gg3d_hbsm"])
def test_uniform_state_has_only_well_and_bulk_energy(name):
    cfg, st = parity.make_case(name)
    y = _np_state(st)
    y["phase"][...] = 0.5
    q = np.zeros_like(y["quat"])
    q[0] = 1.0
    y["quat"] = q
    if y.get("conc") is not None:
        y["conc"][...] = 0.08
    if y.get("temperature") is not None:
        y["temperature"][...] = 0.7
    o = pyoracle.Oracle(cfg)
    if cfg.conc_rhs_form in (2, 3):
        o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
    status, e = o.energy(y)
    assert status == 0
    vol = _volume(cfg)
    # well_func('d') = 16 phi^2 (1-phi)^2 = 1 at phi = 1/2
    assert e[4] == pytest.approx(cfg.phi_well_scale * vol, rel=1e-13)
    assert e[1] == 0.0 and e[3] == 0.0
    # |grad q| = 0 and floor type 'm' (no additive floor, quatenergy.m4:246-250): no orientational part
    assert cfg.grad_floor_type == b"m"
    assert e[2] == 0.0
    assert e[0] == pytest.approx(e[1] + e[2] + e[3] + e[4] + e[5], rel=1e-13)
    if cfg.conc_rhs_form == 2:
        # quadratic KKS at uniform c: c_l, c_a from the closed form, f = A (c - ceq)^2 / Vm
        cl, ca = o.phase_concentrations()
        assert np.ptp(cl) < 1e-15 and np.ptp(ca) < 1e-15
        assert e[5] != 0.0


def test_isotropic_interface_energy_is_the_discrete_dirichlet_form():
    """sum phi (-Lap_h phi) = sum over faces (dphi)^2 / h^2 on a periodic grid (summation by parts)"""
    cfg, st = parity.make_case("gg3d_hbsm")
    y = _np_state(st)
    o = pyoracle.Oracle(cfg)
    o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
    status, e = o.energy(y)
    assert status == 0
    phi = y["phase"].reshape(cfg.n[2], cfg.n[1], cfg.n[0])
    w = cfg.dx[0] * cfg.dx[1] * cfg.dx[2]
    s = 0.0
    for ax, d in ((2, 0), (1, 1), (0, 2)):
        s += (((np.roll(phi, -1, ax) - phi) / cfg.dx[d]) ** 2).sum()
    ref = 0.5 * cfg.epsilon_phase ** 2 * s * w
    assert e[1] == pytest.approx(ref, rel=1e-11)


def test_pfhub1a_energy_closed_forms():
    cfg, st = parity.make_case("pfhub1a")
    o = pyoracle.Oracle(cfg)
    vol = _volume(cfg)
    y = _np_state(st)
    y["conc"][...] = cfg.ch_ca
    assert o.energy(y)[1][0] == 0.0
    y["conc"][...] = 0.5 * (cfg.ch_ca + cfg.ch_cb)
    half = 0.5 * (cfg.ch_cb - cfg.ch_ca)
    assert o.energy(y)[1][0] == pytest.approx(cfg.ch_well_scale * half ** 4 * vol, rel=1e-13)
    # gradient part of a plane wave: kappa/2 * mean over the two faces of (dc/h)^2
    nx, ny = cfg.n[0], cfg.n[1]
    x = (np.arange(nx) + 0.5) * cfg.dx[0]
    c = 0.5 + 0.05 * np.cos(2 * np.pi * 3 * x / (nx * cfg.dx[0]))
    y["conc"] = np.ascontiguousarray(np.broadcast_to(c, (1, ny, nx)).copy())
    e = o.energy(y)[1]
    g = (np.roll(c, -1) - c) / cfg.dx[0]
    grad = 0.5 * cfg.ch_kappa * (g ** 2).sum() * ny * cfg.dx[0] * cfg.dx[1]
    assert e[1] == pytest.approx(grad, rel=1e-12)
    well = (cfg.ch_well_scale * (c - cfg.ch_ca) ** 2 * (cfg.ch_cb - c) ** 2).sum() * ny * cfg.dx[0] * cfg.dx[1]
    assert e[4] == pytest.approx(well, rel=1e-12)
    assert e[0] == pytest.approx(e[1] + e[4], rel=1e-14)


def test_explicit_cahn_hilliard_steps_decrease_the_energy():
    cfg, st = parity.make_case("pfhub1a")
    _, energies = parity.oracle_trajectory(cfg, st, parity.TRAJ_DT["pfhub1a"], 60, energy_every=10)
    f = np.array([e[0] for e in energies])
    assert np.all(np.diff(f) < 0.0)


# ---- QuatModel::printScalarDiagnostics ------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["dendrite2d", "auni3d", "pfhub1a"])
def test_scalar_diagnostics_restatement(name):
    """volume of solid = L1 norm of phi times the cell volume, integral / max concentration, Cex =
    (int |c phi| - c0 int |phi|) / int c, temperature extrema / average, thermal energy = -L int phi + cp int T
    (QuatModel.cc:2543-2690, 5106-5180, 5373-5392) against numpy"""
    import parity
    from oracle import pyoracle
    cfg, st = parity.make_case(name)
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    d = o.scalar_diagnostics(y)
    o.close()
    dv = float(np.prod([cfg.dx[a] for a in range(cfg.ndim)]))
    ncell = cfg.n[0] * cfg.n[1] * (cfg.n[2] if cfg.ndim == 3 else 1)
    vol = dv * ncell
    assert d["volume"] == pytest.approx(vol, rel=1e-14)
    if cfg.with_phase:
        assert d["volume_solid"] == pytest.approx(np.abs(y["phase"]).sum() * dv, rel=1e-12)
        assert d["solid_fraction"] == pytest.approx(np.abs(y["phase"]).mean(), rel=1e-12)
    else:
        assert d["solid_fraction"] == 1.0
    if cfg.with_concentration:
        c = y["conc"]
        phi = y["phase"] if cfg.with_phase else np.ones_like(c)
        assert d["integral_concentration"] == pytest.approx(c.sum() * dv, rel=1e-12)
        assert d["max_concentration"] == c.max()
        cphi = np.abs(c * phi).sum() * dv
        assert d["integral_phase_concentration"] == pytest.approx(cphi, rel=1e-12)
        expect = (cphi - c.sum() * dv / vol * d["volume_solid"]) / (c.sum() * dv)
        assert d["cex"] == pytest.approx(expect, rel=1e-9, abs=1e-13)
    if cfg.with_unsteady_temperature:
        T = y["temperature"]
        assert d["min_temperature"] == T.min() and d["max_temperature"] == T.max()
        assert d["average_temperature"] == pytest.approx(T.mean(), rel=1e-12)
        assert d["thermal_energy"] == pytest.approx(-cfg.latent_heat * y["phase"].sum() * dv + cfg.cp * T.sum() * dv,
                                                    rel=1e-11)
    else:
        assert d["min_temperature"] == d["max_temperature"] == d["average_temperature"] == cfg.T_uniform
