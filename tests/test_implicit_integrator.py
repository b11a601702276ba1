"""Host logic of the implicit integrator (ampe_b200/host/ImplicitIntegrator.h) on
the CPU: (1) a toy vector backend with closed-form discrete solutions pins the BDF1/BDF2 coefficients,
the Newton iteration and the matrix-free GMRES; (2) the same template driven by the oracle's RHS is
checked against small-step explicit trajectories (observed order 1 and 2), mass conservation and the
quaternion constraint.  The device backend is compared with (2) in tests/test_gpu_widening_z_implicit.py."""
import os
import subprocess

import numpy as np
import pytest

import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def toy(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("toy") / "implicit_toy")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "implicit_toy.cpp")])
    out = subprocess.check_output([exe]).decode()
    return {k: float(v) for k, v in (line.split() for line in out.strip().splitlines())}


def test_bdf_coefficients_against_closed_form(toy):
    """periodic diffusion, three Fourier modes: BDF1 amplification 1/(1 - lambda h) per step, BDF2 two-step
    recurrence with a BDF1 start, at a step 6x the explicit stability limit"""
    for order in (1, 2):
        assert toy["diffusion_bdf%d_rc" % order] == 0
        assert toy["diffusion_bdf%d_err" % order] < 1e-9
        assert toy["diffusion_bdf%d_jtimes_fd1" % order] == toy["diffusion_bdf%d_linear_iterations" % order] > 0
        assert toy["diffusion_bdf%d_rhs_fd0" % order] >= 26  # predictor + at least one residual per step


def test_newton_solves_the_nonlinear_step(toy):
    assert toy["cubic_rc"] == 0
    assert toy["cubic_residual"] < 1e-10  # y1 - y0 + h y1^3 = 0
    assert toy["cubic_newton_iterations"] >= 2
    assert abs(toy["cubic_bdf1_observed_order"] - 1.0) < 0.1
    assert abs(toy["cubic_bdf2_observed_order"] - 2.0) < 0.15


def test_failures_are_reported(toy):
    assert toy["starved_newton_rc"] == -20  # IMPLICIT_ENEWTON
    assert toy["bad_step_rc"] == -1  # IMPLICIT_EINVAL


def _implicit(name, dt, nsteps, **kw):
    from oracle import pyoracle
    cfg, st = parity.make_case(name)
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    if cfg.conc_rhs_form in (2, 3):
        o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
    rot = parity.random_rotations(cfg) if cfg.symmetry_aware else None
    if rot is not None:
        o.set_rotations(rot)
    rc, stats = o.integrate_implicit(y, dt, nsteps, **kw)
    o.close()
    return cfg, st, y, rc, stats


def test_pfhub1a_observed_order_and_mass():
    """Cahn-Hilliard (C1): against a Heun trajectory with a 16x smaller step the BDF1 error halves and the
    BDF2 error quarters when the step is halved; total composition is conserved"""
    dt = parity.TRAJ_DT["pfhub1a"]
    cfg, st = parity.make_case("pfhub1a")
    yref, _ = parity.oracle_trajectory(cfg, st, dt / 4, 160, scheme=1)
    tight = dict(rtol=1e-9, atol=1e-11, max_krylov=30, max_newton=6)
    errs = {}
    for order in (1, 2):
        for mult in (4, 2):
            _, _, y, rc, stats = _implicit("pfhub1a", dt * mult, 40 // mult, order=order, **tight)
            assert rc == 0
            errs[order, mult] = np.abs(y["conc"] - yref["conc"]).max()
            assert abs(y["conc"].sum() - st["conc"].numpy().sum()) < 1e-10
    assert 1.7 < errs[1, 4] / errs[1, 2] < 2.3
    assert 3.4 < errs[2, 4] / errs[2, 2] < 4.6
    assert errs[2, 2] < 0.2 * errs[1, 2]


@pytest.mark.parametrize("name,mult", [("dendrite2d", 50), ("auni2d", 20)])
def test_stiff_steps_with_the_full_model(name, mult):
    """steps of 10x (4x) the explicit stability limit: Newton converges, GMRES works through fd_flag = 1
    products, quaternions stay on the unit sphere (projection + normalizeQuat)"""
    dt = parity.TRAJ_DT[name] * mult
    cfg, st, y, rc, stats = _implicit(name, dt, 3, order=2, rtol=1e-8, atol=1e-10, max_krylov=30, max_newton=8)
    assert rc == 0, stats
    assert stats["steps"] == 3 and stats["projections"] == 3
    assert stats["jtimes_evals"] == stats["linear_iterations"] > 0
    q = y["quat"].reshape(cfg.qlen, -1)
    assert np.abs((q * q).sum(0) - 1.0).max() < 1e-14
    for k in ("phase", "quat", "conc", "temperature"):
        if y.get(k) is not None:
            assert np.isfinite(y[k]).all()
    assert np.abs(y["phase"] - st["phase"].numpy()).max() > 0.0


def test_default_tolerances_are_the_reference_defaults():
    """atol 3e-4, rtol = atol*1e-2, max_order 2, Krylov dimension 5 (QuatIntegrator.cc:285-301)"""
    text = open(os.path.join(ROOT, "ampe_b200", "host", "ImplicitIntegrator.h")).read()
    for frag in ("order = 2", "rtol = 3.e-6, atol = 3.e-4", "max_krylov_dimension = 5"):
        assert frag in text
    _, _, _, rc, stats = _implicit("pfhub1a", parity.TRAJ_DT["pfhub1a"] * 20, 5)
    assert rc == 0 and stats["steps"] == 5


# ---- advanceTo: variable steps with the local error test ------------------------------------------------------
def test_adaptive_controller_on_closed_form_problems(toy):
    """semi-discrete diffusion (decay rates 27 .. 1290) against its exact solution: the run lands on tend, the
    step grows by orders of magnitude as the fast modes die, the work scales like tol^(-1/3) (second order) and
    the global error follows the tolerance (error-per-step control: ~tol^(2/3)); failure codes are reported"""
    for k in range(3):
        assert toy["adaptive%d_rc" % k] == 0
        assert toy["adaptive%d_t_reached_err" % k] == 0.0
        assert toy["adaptive%d_step_growth" % k] > 100.0
        assert toy["adaptive%d_error_test_failures" % k] <= 3
    assert toy["adaptive0_err_over_tol"] < 5 and toy["adaptive1_err_over_tol"] < 20 and toy["adaptive2_err_over_tol"] < 80
    # 100x tighter tolerance: 100^(1/3) = 4.6x the steps for a second-order method
    assert 3.0 < toy["adaptive1_steps"] / toy["adaptive0_steps"] < 6.0
    assert 3.0 < toy["adaptive2_steps"] / toy["adaptive1_steps"] < 6.0
    assert toy["adaptive_cubic_rc"] == 0 and toy["adaptive_cubic_err"] < 5e-5
    assert toy["adaptive_too_much_work_rc"] == -22
    assert toy["adaptive_hmin_rc"] == -23
    assert toy["adaptive_bad_interval_rc"] == -1


def test_adaptive_pfhub1a_against_a_fine_trajectory():
    """Cahn-Hilliard (C1) from t = 0 to 200 explicit steps' worth of time: the adaptive run agrees with a Heun
    trajectory of 4x smaller steps to its tolerance, conserves the total composition, and needs far fewer steps
    than the explicit integrator because the step grows while the spinodal structure coarsens"""
    from oracle import pyoracle
    dt = parity.TRAJ_DT["pfhub1a"]
    cfg, st = parity.make_case("pfhub1a")
    tend = 200 * dt
    yref, _ = parity.oracle_trajectory(cfg, st, dt / 4, 800, scheme=1)
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    rc, stats = o.integrate_adaptive(y, tend, dt, rtol=1e-6, atol=1e-8, max_krylov=30, max_newton=4, max_steps=2000)
    o.close()
    assert rc == 0, stats
    assert abs(stats["t_reached"] - tend) < 1e-12 * tend
    assert stats["steps"] < 120 and stats["largest_step"] > 3 * stats["smallest_step"], stats
    assert np.abs(y["conc"] - yref["conc"]).max() < 2e-5, stats
    assert abs(y["conc"].sum() - st["conc"].numpy().sum()) < 1e-10


def test_adaptive_full_model_with_preconditioner():
    """GG3D_HBSM over 120 explicit steps' worth of time with AMPE's default tolerances, preconditioned: finishes,
    quaternions stay unit, composition is conserved, steps beyond the explicit limit are taken"""
    from oracle import pyoracle
    name = "gg3d_hbsm"
    dt = parity.TRAJ_DT[name]
    cfg, st = parity.make_case(name)
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    o.set_preconditioner(2)
    rc, stats = o.integrate_adaptive(y, 120 * dt, dt, max_krylov=10, max_newton=4)
    o.close()
    assert rc == 0, stats
    assert stats["largest_step"] > 6 * dt, stats  # the explicit stability limit is ~5 dt
    q = y["quat"].reshape(cfg.qlen, -1)
    assert np.abs((q * q).sum(0) - 1.0).max() < 1e-14
    assert abs(y["conc"].sum() - st["conc"].numpy().sum()) < 1e-9 * abs(st["conc"].numpy().sum())


def test_adaptive_pfhub1a_free_energy_decreases():
    """PFHub benchmark 1a is a gradient flow: along a variable-step implicit run over 2000 explicit steps' worth
    of time the free energy sampled 20 times never increases, while the step grows by more than 10x"""
    from oracle import pyoracle
    dt = parity.TRAJ_DT["pfhub1a"]
    cfg, st = parity.make_case("pfhub1a")
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    t, h, energies, hmax = 0.0, dt, [o.energy(y)[1][0]], 0.0
    for k in range(20):
        rc, s = o.integrate_adaptive(y, t + 100 * dt, h, t0=t, rtol=1e-6, atol=1e-8, max_krylov=30, max_newton=4,
                                     max_steps=2000)
        assert rc == 0, s
        t, h, hmax = s["t_reached"], s["last_step"], max(hmax, s["largest_step"])
        energies.append(o.energy(y)[1][0])
    o.close()
    assert all(b <= a + 1e-12 * abs(a) for a, b in zip(energies, energies[1:])), energies
    assert energies[-1] < 0.97 * energies[0]  # 319 -> 301 over this interval
    assert hmax > 10 * dt
