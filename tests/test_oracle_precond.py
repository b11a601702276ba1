"""Block preconditioners on the CPU.

(1) The product's per-cell multigrid arithmetic (ampe_b200/csrc/mg_cell.h, looped on the host by
    oracle/precond.cc part 2) applies the SAME operator as the reference's routines restated on their own
    layouts (efo_compfluxvardc + efo_compresvarsca, set_j_ij + set_stencil), and both reproduce the closed-form
    eigenvalues of the constant-coefficient operator.
(2) The solve: the reference hands its single level to hypre PFMG (absent from its tree), so there is no solver
    to restate -- the V-cycle is judged by the residual the RESTATED operator measures (contraction per cycle,
    odd extents, jumping coefficients).
(3) CVSpgmrPrecondSet / CVSpgmrPrecondSolve on the five configurations: operator identity with the context's own
    coefficients, residual after the default two cycles, and what the preconditioner does to the Newton-Krylov
    iteration of the implicit integrator (fewer Krylov vectors, same trajectory).
The device solver is compared with this host loop in tests/test_gpu_widening_zzz_precond.py."""
import numpy as np
import pytest

import parity
from oracle import pyoracle


def _ghosted(a, ng, ndim):
    """periodic ghost fill of a (nz, ny, nx) cell array in the first ndim directions"""
    pad = [(ng, ng) if (3 - 1 - ax) < ndim else (0, 0) for ax in range(3)]
    return np.ascontiguousarray(np.pad(a, pad, mode="wrap"))


def _side_from_lower(low, axis_np):
    """SideData array of one direction from its ghost-0 lower-face values (periodic: the extra upper
    face repeats face 0)"""
    first = np.take(low, [0], axis=axis_np)
    return np.ascontiguousarray(np.concatenate([low, first], axis=axis_np))


def _random_elliptic(n, seed, jump=1.0):
    rng = np.random.default_rng(seed)
    ndim = len(n)
    shape = (n[2] if ndim == 3 else 1, n[1], n[0])
    m = 0.5 + rng.random(shape)
    c = 1.0 + rng.random(shape)
    lows = [-(0.2 + rng.random(shape)) * np.where(rng.random(shape) < 0.3, jump, 1.0) for _ in range(ndim)]
    d = [_side_from_lower(lows[a], 2 - a) for a in range(ndim)]
    return shape, m, c, lows, d


@pytest.mark.parametrize("n,dx", [((12, 10), (0.3, 0.2)), ((8, 6, 10), (0.5, 0.4, 0.25)), ((9, 7), (1.0, 1.0))])
def test_scalar_operator_is_the_reference_operator(n, dx):
    """M div(D grad u) + C u with variable M, C, D: product arithmetic == restated efo_* routines"""
    ndim = len(n)
    shape, m, c, lows, d = _random_elliptic(n, 3)
    rng = np.random.default_rng(4)
    u = rng.standard_normal(shape)
    mg = pyoracle.HostMG(n, dx)
    mg_m, mg_c = _ghosted(m, 1, ndim), _ghosted(c, 2, ndim)
    mg.set_elliptic(m=mg_m, ngm=1, c=mg_c, ngc=2, d=d, ngd=0, d_scale=1.0)
    ref = pyoracle.elliptic_apply(n, dx, mg_m, 1, mg_c, 2, d, u)
    got = mg.apply(u)
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()
    # the scale factor and the second diffusion array (EBS: D_l + D_a, times -gamma)
    mg.set_elliptic(m=None, m_const=0.7, c=None, c_const=1.0, d=d, d2=d, ngd=0, d_scale=-0.25)
    d_eff = [-0.25 * (x + x) for x in d]
    ones = np.ones(shape)
    ref2 = pyoracle.elliptic_apply(n, dx, 0.7 * ones, 0, ones, 0, d_eff, u)
    assert np.abs(mg.apply(u) - ref2).max() <= 1e-13 * np.abs(ref2).max()


@pytest.mark.parametrize("n,dx", [((12, 10), (0.3, 0.2)), ((8, 6, 10), (0.5, 0.4, 0.25))])
def test_quaternion_operator_is_the_reference_stencil(n, dx):
    """w + gamma sqrt(m) div(fc grad(sqrt(m) w)): product arithmetic == restated set_j_ij + set_stencil"""
    ndim = len(n)
    rng = np.random.default_rng(5)
    shape = (n[2] if ndim == 3 else 1, n[1], n[0])
    mob = 0.1 + rng.random(shape)
    lows = [-(0.2 + rng.random(shape)) for _ in range(ndim)]
    fc = [_side_from_lower(lows[a], 2 - a) for a in range(ndim)]
    w = rng.standard_normal(shape)
    gamma = 0.37
    mg = pyoracle.HostMG(n, dx, with_s=True)
    mob_g = _ghosted(mob, 1, ndim)
    mg.set_quat(gamma, mob_g, 1, fc, 0)
    ref = pyoracle.quat_stencil_apply(n, dx, gamma, np.sqrt(mob_g), 1, fc, w)
    got = mg.apply(w)
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()
    # symmetric matrix (what set_symmetric_stencil relies on): <x, A y> == <A x, y>
    x = rng.standard_normal(shape)
    assert abs((x * mg.apply(w)).sum() - (mg.apply(x) * w).sum()) < 1e-11 * np.abs(got).sum()


def test_constant_coefficient_eigenvalues():
    """Fourier modes are eigenvectors: lambda = C + M D sum_a (2 cos(2 pi k_a / n_a) - 2) / h_a^2"""
    n, dx = (16, 12), (0.5, 0.25)
    M, Cc, D = 1.3, 1.0, -0.04
    mg = pyoracle.HostMG(n, dx)
    mg.set_elliptic(m_const=M, c_const=Cc, d_const=D)
    jj, ii = np.meshgrid(np.arange(n[1]), np.arange(n[0]), indexing="ij")
    for k in ((0, 0), (1, 0), (3, 5), (8, 6)):
        u = np.cos(2 * np.pi * (k[0] * ii / n[0] + k[1] * jj / n[1]))[None]
        lam = Cc + M * D * sum((2 * np.cos(2 * np.pi * k[a] / n[a]) - 2) / dx[a] ** 2 for a in range(2))
        assert np.abs(mg.apply(u) - lam * u).max() < 1e-12 * max(1.0, abs(lam))
        # and the V-cycles invert it: one mode in, the same mode out, scaled by 1 / lambda
        z = mg.solve(u, ncycles=10)
        assert np.abs(z - u / lam).max() < 1e-9


def _contraction(mg, apply_ref, rhs, cycles, symmetrized=False):
    hist = []
    for nc in cycles:
        z = mg.solve(rhs, ncycles=nc, symmetrized=symmetrized)
        hist.append(np.linalg.norm(rhs - apply_ref(z)) / np.linalg.norm(rhs))
    return hist


@pytest.mark.parametrize("n,dx,stiff", [((64, 48), (1.0, 1.0), 400.0), ((32, 16, 24), (1.0, 1.0, 1.0), 50.0)])
def test_vcycle_contracts_the_reference_residual(n, dx, stiff):
    """diffusion number gamma D / h^2 >> 1 (an essentially singular-perturbed Poisson problem): the residual
    measured by the RESTATED operator drops by >= 4x per cycle (cell-to-cell random coefficients, contrast 6) and keeps dropping"""
    ndim = len(n)
    shape, m, c, lows, d = _random_elliptic(n, 8)
    m[:] = 1.0
    c[:] = 1.0
    d = [stiff * x for x in d]
    rng = np.random.default_rng(9)
    rhs = rng.standard_normal(shape)
    mg = pyoracle.HostMG(n, dx)
    assert mg.num_levels() >= 4
    mg.set_elliptic(m=m, ngm=0, c=c, ngc=0, d=d, ngd=0)
    hist = _contraction(mg, lambda z: pyoracle.elliptic_apply(n, dx, m, 0, c, 0, d, z), rhs, (1, 2, 3, 4, 14))
    assert hist[0] < 0.25
    for a, b in zip(hist[:3], hist[1:4]):
        assert b < 0.25 * a, hist
    assert hist[-1] < 1e-7, hist


def test_coarse_levels_are_rediscretised_means():
    n, dx = (16, 8), (1.0, 0.5)
    shape, m, c, lows, d = _random_elliptic(n, 12)
    mg = pyoracle.HostMG(n, dx)
    mg.set_elliptic(m=m, ngm=0, c=c, ngc=0, d=d, ngd=0)
    assert [mg.level_extents(l) for l in range(mg.num_levels())] == [[16, 8, 1], [8, 4, 1], [4, 2, 1]]
    c1 = mg.level_array(1, 0)[0]
    assert np.allclose(c1, c[0].reshape(4, 2, 8, 2).mean(axis=(1, 3)), rtol=0, atol=1e-15)
    # x faces: mean of the two fine faces stacked in y, / h^2 with h doubled
    d0 = mg.level_array(0, 3)[0]
    assert np.allclose(d0, lows[0][0] / dx[0] ** 2, rtol=1e-15)
    d0c = mg.level_array(1, 3)[0]
    expect = 0.25 * d0[:, ::2].reshape(4, 2, 8).mean(axis=1)
    assert np.allclose(d0c, expect, rtol=1e-14)
    d1c = mg.level_array(1, 4)[0]
    d1 = mg.level_array(0, 4)[0]
    assert np.allclose(d1c, 0.25 * d1[::2, :].reshape(4, 8, 2).mean(axis=2), rtol=1e-14)


def test_odd_extents_fall_back_to_jacobi_and_jumps_still_converge():
    """an odd extent cannot be two-coloured across the periodic wrap: single level, damped Jacobi; a 1e3
    diffusivity ratio across a diffuse interface (solid / liquid) does not break the cycle"""
    n, dx = (15, 9), (1.0, 1.0)
    shape, m, c, lows, d = _random_elliptic(n, 21)
    rng = np.random.default_rng(22)
    rhs = rng.standard_normal(shape)
    mg = pyoracle.HostMG(n, dx)
    assert mg.num_levels() == 1
    mg.set_sweeps(1, 1, 60)
    mg.set_elliptic(m=m, ngm=0, c=c, ngc=0, d=d, ngd=0)
    z = mg.solve(rhs, ncycles=4)
    r = rhs - pyoracle.elliptic_apply(n, dx, m, 0, c, 0, d, z)
    assert np.linalg.norm(r) < 1e-6 * np.linalg.norm(rhs)
    # a solid disc in liquid, interface 4 cells wide, diffusivity ratio 1e3 (face value = mean of the cells)
    n = (64, 64)
    jj, ii = np.meshgrid(np.arange(n[1]), np.arange(n[0]), indexing="ij")
    phi = 0.5 * (1.0 - np.tanh((np.hypot(ii - 31.5, jj - 31.5) - 18.0) / 2.0))[None]
    dcell = -40.0 * (1.0e-3 * phi + (1.0 - phi))
    lows = [0.5 * (dcell + np.roll(dcell, 1, axis=2 - a)) for a in range(2)]
    d = [_side_from_lower(lows[a], 2 - a) for a in range(2)]
    ones = np.ones_like(phi)
    rhs = rng.standard_normal(phi.shape)
    mg = pyoracle.HostMG(n, dx)
    mg.set_elliptic(m_const=1.0, c_const=1.0, d=d, ngd=0)
    hist = _contraction(mg, lambda z: pyoracle.elliptic_apply(n, dx, ones, 0, ones, 0, d, z), rhs, (1, 2, 3, 12))
    assert hist[0] < 0.3 and hist[1] < 0.3 * hist[0] and hist[2] < 0.3 * hist[1] and hist[3] < 1e-6, hist


# ---- CVSpgmrPrecondSet / CVSpgmrPrecondSolve on the configurations -----------------------------------------
BLOCKS = {"phase": 0, "quat": 1, "conc": 2, "temperature": 3}


def _context(name):
    cfg, st = parity.make_case(name)
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    if cfg.conc_rhs_form in (2, 3):
        o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
    if cfg.symmetry_aware:
        o.set_rotations(parity.random_rotations(cfg))
    status, _ = o.eval(0.0, y, fd_flag=0)
    assert status == 0
    return cfg, y, o


def _evolved(cfg):
    out = []
    if cfg.with_phase:
        out.append("phase")
    if cfg.evolve_quat:
        out.append("quat")
    if cfg.with_concentration and cfg.conc_rhs_form in (2, 3):
        out.append("conc")
    if cfg.with_unsteady_temperature:
        out.append("temperature")
    return out


@pytest.mark.parametrize("name", ["dendrite2d", "auni2d", "gg3d_hbsm", "auni3d"])
def test_precond_set_and_solve_on_the_configurations(name):
    """gamma = 20x the explicit step: every block's multigrid applies the restated reference operator built
    from the context's own coefficient arrays, and two V-cycles leave < 10 % of the residual (the
    quaternion block through divide/multiplyMobilitySqrt)"""
    cfg, y, o = _context(name)
    gamma = 20 * parity.TRAJ_DT[name]
    assert o.precond_setup(gamma, 2) == 0
    rng = np.random.default_rng(31)
    shape = y["phase"].shape
    r = {k: (None if v is None else rng.standard_normal(v.shape)) for k, v in y.items()}
    rc, z = o.precond_solve(r)
    assert rc == 0
    for k in _evolved(cfg):
        b = BLOCKS[k]
        mg = o.precond_block(b)
        u = rng.standard_normal(shape)
        ref = o.precond_apply(b, u)
        assert np.abs(mg.apply(u) - ref).max() <= 1e-12 * np.abs(ref).max(), k
        if k == "quat":
            # QuatSysSolver::solveSystem: the level solver sees rhs / sqrt(m) and returns w, z = sqrt(m) w
            s = mg.level_array(0, 2).reshape(shape)
            for m in range(cfg.qlen):
                res = r[k][m] / s - o.precond_apply(b, z[k][m] / s)
                assert np.linalg.norm(res) < 0.1 * np.linalg.norm(r[k][m] / s), (k, m)
        else:
            res = r[k] - o.precond_apply(b, z[k])
            assert np.linalg.norm(res) < 0.1 * np.linalg.norm(r[k]), k
    o.close()


def test_phase_block_coefficients_follow_phasefacops():
    """C = 1 + gamma M w g''(phi), g'' = 32 (1 + 6 phi (phi - 1)); D = -gamma eps^2 / h^2; M = phi_mobility"""
    cfg, y, o = _context("auni2d")
    gamma = 3.0e-9
    o.precond_setup(gamma, 2)
    mg = o.precond_block(0)
    phi = y["phase"].reshape(mg.level_array(0, 0).shape)
    expect = 1.0 + gamma * cfg.phi_mobility * cfg.phi_well_scale * 32.0 * (1.0 + 6.0 * phi * (phi - 1.0))
    assert np.allclose(mg.level_array(0, 0), expect, rtol=1e-14)
    assert np.allclose(mg.level_array(0, 1), cfg.phi_mobility, rtol=0)
    assert np.allclose(mg.level_array(0, 3), -gamma * cfg.epsilon_phase ** 2 / cfg.dx[0] ** 2, rtol=1e-15)
    assert mg.level_array(0, 2) is None
    o.close()


def _implicit(name, dt, nsteps, precond, **kw):
    cfg, st = parity.make_case(name)
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    if cfg.conc_rhs_form in (2, 3):
        o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
    if cfg.symmetry_aware:
        o.set_rotations(parity.random_rotations(cfg))
    if isinstance(precond, tuple):
        o.set_preconditioner(precond[0], dquatdphi=precond[1])
    else:
        o.set_preconditioner(precond)
    rc, stats = o.integrate_implicit(y, dt, nsteps, **kw)
    stats.update(o.precond_stats())
    o.close()
    return cfg, y, rc, stats


@pytest.mark.parametrize("name,mult,ratio", [("dendrite2d", 50, 0.9), ("gg3d_hbsm", 40, 0.4), ("auni2d", 20, 0.7),
                                             ("auni3d", 20, 0.5)])
def test_preconditioner_cuts_the_krylov_work_and_keeps_the_trajectory(name, mult, ratio):
    """stiff steps (20-50x the explicit trajectory step): right-preconditioned GMRES needs fewer Jacobian-vector
    products for the same Newton tolerance (measured: 266 -> 60 for GG3D, 35 -> 13 for AuNi_3D, 20 -> 11 for
    AuNi_2D; only 134 -> 104 for Dendrite2D, whose stiffness sits in the derivative of the singular 1/|grad q|
    diffusivity that the frozen-coefficient blocks -- the reference's too -- do not contain), and the step it
    converges to is the same"""
    dt = parity.TRAJ_DT[name] * mult
    kw = dict(order=2, rtol=1e-8, atol=1e-10, max_krylov=30, max_newton=8)
    cfg, y0, rc0, s0 = _implicit(name, dt, 3, 0, **kw)
    cfg, y1, rc1, s1 = _implicit(name, dt, 3, 2, **kw)
    assert rc0 == 0 and rc1 == 0, (s0, s1)
    assert s0["precond_setups"] == 0 and s1["precond_setups"] == s1["newton_iterations"] > 0
    assert s1["precond_solves"] >= s1["linear_iterations"]
    assert s1["linear_iterations"] < ratio * s0["linear_iterations"], (s0, s1)
    for k in ("phase", "quat", "conc", "temperature"):
        if y0.get(k) is not None:
            scale = max(np.abs(y0[k]).max(), 1e-300)
            assert np.abs(y1[k] - y0[k]).max() < 1e-6 * scale, k


# ---- the dquat/dphi coupling block (precond_has_dquatdphi) ----------------------------------------------------
def _quat_op(d, q, ndim):
    """-sum_a [d_a(i+1) (q(i+1) - q(i)) - d_a(i) (q(i) - q(i-1))] with d = fc / h^2 on lower faces (periodic):
    compute_flux + add_quat_op (rhs = rhs - mobility * divergence) without the mobility
    (2d/quatfacops.m4:173-229, 487-540)"""
    out = np.zeros_like(q)
    for a in range(ndim):
        ax = q.ndim - 1 - a
        flux = d[a][None] * (q - np.roll(q, 1, axis=ax))
        out -= np.roll(flux, -1, axis=ax) - flux
    return out


@pytest.mark.parametrize("name", ["auni2d", "gg3d_hbsm", "auni3d"])
def test_dquatdphi_block_is_the_phase_derivative_of_the_frozen_operator(name):
    """QuatFACOps::multiplyDQuatDPhiBlock restated from quatmobilityderiv / quatdiffusionderiv /
    compute_dquatdphi_face_coef against difference quotients of the level solver's own coefficients set up at
    phi and at phi + eps z:  out = [m'(phi) z] L(fc, q) + sqrt(m) L(fc'[z], q)  (the second term carries
    sqrt(m), not m: d_sqrt_m_id at QuatFACOps.cc:1953, restated as is)"""
    rng = np.random.default_rng(41)
    gamma = 20 * parity.TRAJ_DT[name]
    cfg, y, o = _context(name)
    o.precond_setup(gamma, 2, dquatdphi=True)
    shape = y["phase"].shape
    z = rng.standard_normal(shape)
    got = o.precond_dquatdphi(z).reshape((cfg.qlen,) + shape[-3:])
    mg0 = o.precond_block(1)
    s0 = mg0.level_array(0, 2)
    d0 = [mg0.level_array(0, 3 + a) for a in range(cfg.ndim)]
    eps = 1e-7
    cfg1, st1 = parity.make_case(name)
    y1 = {k: (None if v is None else v.numpy().copy()) for k, v in st1.items()}
    y1["phase"] = y1["phase"] + eps * z
    o1 = pyoracle.Oracle(cfg1)
    if cfg1.conc_rhs_form in (2, 3):
        o1.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
    if cfg1.symmetry_aware:
        o1.set_rotations(parity.random_rotations(cfg1))
    assert o1.eval(0.0, y1, fd_flag=0)[0] == 0
    o1.precond_setup(gamma, 2)
    mg1 = o1.precond_block(1)
    s1 = mg1.level_array(0, 2)
    d1 = [mg1.level_array(0, 3 + a) for a in range(cfg.ndim)]
    q = y["quat"].reshape((cfg.qlen,) + s0.shape)
    dm = (s1 * s1 - s0 * s0) / eps
    dfc = [(d1[a] - d0[a]) / eps for a in range(cfg.ndim)]
    expect = dm[None] * _quat_op(d0, q, cfg.ndim) + s0[None] * _quat_op(dfc, q, cfg.ndim)
    scale = np.abs(expect).max()
    assert scale > 0
    assert np.abs(got.reshape(expect.shape) - expect).max() < 2e-4 * scale
    o.close()
    o1.close()


@pytest.mark.parametrize("name,mult", [("gg3d_hbsm", 40), ("auni3d", 20)])
def test_coupled_preconditioner_in_the_integrator(name, mult):
    """the lower-triangular coupling (z_q sees gamma dF_q/dphi z_phi, QuatIntegrator.cc:3602-3612) keeps the
    trajectory and does not cost Krylov vectors"""
    dt = parity.TRAJ_DT[name] * mult
    kw = dict(order=2, rtol=1e-8, atol=1e-10, max_krylov=30, max_newton=8)
    cfg, y1, rc1, s1 = _implicit(name, dt, 3, 2, **kw)
    cfg, y2, rc2, s2 = _implicit(name, dt, 3, (2, True), **kw)
    assert rc1 == 0 and rc2 == 0, (s1, s2)
    assert s2["linear_iterations"] <= s1["linear_iterations"] + 2, (s1, s2)
    for k in ("phase", "quat", "conc"):
        if y1.get(k) is not None:
            assert np.abs(y2[k] - y1[k]).max() < 1e-6 * max(np.abs(y1[k]).max(), 1e-300), k


# ---- the reference's own solver tests, mapped onto periodic domains by reflection ----------------------------
def facpoisson_case(ndim):
    """tests/testFACPoisson.cc (FACPoisson.cc:230-231 D = -1, C = 5; setexactandrhs{2,3}d, hyprepoisson.m4:40-75):
    exact = 1 + prod sin(pi x_a) on [0,1]^ndim with value 1 on every boundary, 32^2 / 16^3 cells.  (u - 1) is odd
    about every boundary face, so the periodic problem on [0,2]^ndim with twice the cells has the same discrete
    equations: SAMRAI's Dirichlet ghost value 2 - u_interior IS the periodic image."""
    nc = 32 if ndim == 2 else 16
    h = 1.0 / nc
    pi = 3.141592654  # the reference's literal
    x = (np.arange(2 * nc) + 0.5) * h
    grids = np.meshgrid(*([x] * ndim), indexing="ij")  # axis order irrelevant: symmetric in the coordinates
    sinsin = np.ones_like(grids[0])
    for g in grids:
        sinsin = sinsin * np.sin(pi * g)
    exact = 1.0 + sinsin
    rhs = 5.0 * exact + ndim * pi * pi * sinsin
    shape = (1,) + exact.shape if ndim == 2 else exact.shape
    return [2 * nc] * ndim, [h] * ndim, exact.reshape(shape), rhs.reshape(shape), nc


@pytest.mark.parametrize("ndim", [2, 3])
def test_reference_kat_facpoisson(ndim):
    """acceptance of testFACPoisson.cc:274: max |computed - exact| < 1e-2 after at most 10 cycles"""
    n, dx, exact, rhs, nc = facpoisson_case(ndim)
    mg = pyoracle.HostMG(n, dx)
    mg.set_elliptic(m_const=1.0, c_const=5.0, d_const=-1.0)
    z = mg.solve(rhs, ncycles=10)
    sub = (slice(0, nc if ndim == 3 else 1),) + (slice(0, nc),) * 2
    assert np.abs(z - exact)[sub].max() < 1.0e-2
    assert np.abs(z - exact).max() < (3e-3 if ndim == 2 else 1.2e-2)  # second-order discretisation error
    ones = np.ones_like(rhs)
    side = [np.full(tuple(s + (1 if ax == 2 - a else 0) for ax, s in enumerate(rhs.shape)), -1.0) for a in range(ndim)]
    res = rhs - pyoracle.elliptic_apply(n, dx, ones, 0, 5.0 * ones, 0, side, z)
    # fac_solver.residual_tol = 1e-8 within max_cycles = 10 (FACPoisson/2d.input); 3D contracts by ~0.2 per cycle
    assert np.linalg.norm(res) < (1e-8 if ndim == 2 else 1e-6) * np.linalg.norm(rhs)


def phasefac_case():
    """tests/testPhaseFAC.cc:183-192 + PhaseFAC/2d.input: epsilon 0.1, well scale 0.1, mobility 10, gamma 0.1,
    delta = epsilon / sqrt(32 w); exact = (1 + tanh(x / 2 delta)) / 2 on x in [-1,1] (32 cells), value 0 at x_lo,
    slope 0 at x_up, periodic in y; rhs from phasesetexactandrhs2d (2d/phase.m4:5-49).  Odd reflection about
    x = -1 and even reflection about x = +1 give a period of 8 (128 cells); the coefficient C(phi_exact) is a
    given field and is reflected evenly."""
    eps, w, mob, gamma = 0.1, 0.1, 10.0, 0.1
    delta = eps / np.sqrt(32.0 * w)
    nc, h = 32, 2.0 / 32
    x = -1.0 + h * (np.arange(nc) + 0.5)
    t = 0.5 * (1.0 + np.tanh(0.5 / delta * x))
    f1 = 32.0 * gamma * mob * w
    rhs = -f1 * t * (1.0 - t) * (1.0 - 2.0 * t) + t + f1 * (1.0 + 6.0 * t * (t - 1.0)) * t
    even = lambda a: np.concatenate([a, a[::-1]])            # [-1,1] -> [-1,3]: even about +1
    full = lambda a, s: np.concatenate([even(a), s * even(a)[::-1]])  # -> [-1,7]: odd (s=-1) / even (s=+1) about -1 == 7
    phi_coef = full(t, +1.0)
    return dict(eps=eps, w=w, mob=mob, gamma=gamma, nc=nc, h=h, exact=full(t, -1.0), rhs=full(rhs, -1.0),
                phi_coef=phi_coef)


def test_reference_kat_phasefac():
    """acceptance of testPhaseFAC.cc:60: max |computed - exact| < 1e-2 (the interface is one cell wide)"""
    k = phasefac_case()
    ny = 8
    n, dx = (4 * k["nc"], ny), (k["h"], k["h"])
    tile = lambda a: np.ascontiguousarray(np.tile(a, (1, ny, 1)))
    exact, rhs, phi = tile(k["exact"]), tile(k["rhs"]), tile(k["phi_coef"])
    c = 1.0 + k["gamma"] * k["mob"] * k["w"] * 32.0 * (1.0 + 6.0 * phi * (phi - 1.0))  # PhaseFACOps::setC
    mg = pyoracle.HostMG(n, dx)
    mg.set_elliptic(m_const=k["mob"], c=c, ngc=0, d_const=-k["gamma"] * k["eps"] ** 2)
    z = mg.solve(rhs, ncycles=10)
    err = np.abs(z - exact)[0, :, :k["nc"]].max()
    assert err < 1.0e-2, err


# ---- fused red-black sweep (one pass over tiles, halo of two, ping-pong) ---------------------------------------
@pytest.mark.parametrize("n,dx,quat", [((128, 96), (0.3, 0.2), False), ((64, 16), (1.0, 1.0), False),
                                       ((32, 24, 16), (0.5, 0.4, 0.25), False), ((64, 8, 8), (1.0, 1.0, 1.0), False),
                                       ((96, 80), (0.3, 0.2), True), ((32, 16, 24), (1.0, 1.0, 1.0), True)])
def test_fused_red_black_pass_is_bit_identical(n, dx, quat):
    """mg_rb_tile_pass (what the device runs with one block per tile) == the two colour half-sweeps, bit for
    bit, on every level it applies to -- including tiles as wide as the level (halo cells alias interior cells
    through the periodic wrap) and the quaternion block's column scale"""
    ndim = len(n)
    rng = np.random.default_rng(51)
    shape = (n[2] if ndim == 3 else 1, n[1], n[0])
    rhs = rng.standard_normal(shape)
    outs = []
    for fused in (False, True):
        if quat:
            mob = 0.1 + np.random.default_rng(52).random(shape)
            fc = [_side_from_lower(-(5.0 + 20.0 * np.random.default_rng(53 + a).random(shape)), 2 - a) for a in range(ndim)]
            mg = pyoracle.HostMG(n, dx, with_s=True)
            mg.set_quat(0.37, _ghosted(mob, 1, ndim), 1, fc, 0)
        else:
            _, m, c, lows, d = _random_elliptic(n, 54)
            mg = pyoracle.HostMG(n, dx)
            mg.set_elliptic(m=m, ngm=0, c=c, ngc=0, d=[20.0 * x for x in d], ngd=0)
        mg.set_sweeps(2, 1, 8)
        if fused:
            assert mg.set_fused(True, min_cells=256) >= 1
        outs.append(mg.solve(rhs, ncycles=3, symmetrized=quat))
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("name,mult", [("gg3d_hbsm", 40), ("dendrite2d", 50)])
def test_left_preconditioning_like_the_reference(name, mult):
    """PREC_LEFT (QuatIntegrator.cc:1583): GMRES on P A x = P b stops on the preconditioned residual; the steps it
    converges to are those of the right-preconditioned run (1e-8), with no more Krylov vectors"""
    dt = parity.TRAJ_DT[name] * mult
    kw = dict(order=2, rtol=1e-8, atol=1e-10, max_krylov=30, max_newton=8)
    runs = {}
    for left in (False, True):
        cfg, st = parity.make_case(name)
        y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
        o = pyoracle.Oracle(cfg)
        if cfg.conc_rhs_form in (2, 3):
            o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
        o.set_preconditioner(2, left=left)
        rc, stats = o.integrate_implicit(y, dt, 3, **kw)
        stats.update(o.precond_stats())
        o.close()
        assert rc == 0, stats
        runs[left] = (y, stats)
    assert runs[True][1]["linear_iterations"] <= runs[False][1]["linear_iterations"]
    # one extra solve per linear system (P b) instead of one per solution (P sum c_i v_i)
    assert runs[True][1]["precond_solves"] >= runs[True][1]["linear_iterations"]
    for k, v in runs[False][0].items():
        if v is not None:
            assert np.abs(runs[True][0][k] - v).max() < 1e-8 * max(np.abs(v).max(), 1e-300), k


@pytest.mark.parametrize("n,dx,fused", [((96, 80), (0.3, 0.2), False), ((96, 80), (0.3, 0.2), True),
                                        ((32, 16, 24), (1.0, 1.0, 1.0), True)])
def test_components_solved_together_equal_separate_solves(n, dx, fused):
    """the qlen components of the quaternion block share one matrix: one solver with ncomp components (every
    pass updates all of them, coefficients read once) == one solve per component, bit for bit"""
    ndim = len(n)
    shape = (n[2] if ndim == 3 else 1, n[1], n[0])
    mob = 0.1 + np.random.default_rng(61).random(shape)
    fc = [_side_from_lower(-(5.0 + 20.0 * np.random.default_rng(62 + a).random(shape)), 2 - a) for a in range(ndim)]
    rhs = np.random.default_rng(63).standard_normal((4,) + shape)

    def solver(ncomp):
        mg = pyoracle.HostMG(n, dx, with_s=True, ncomp=ncomp)
        mg.set_quat(0.37, _ghosted(mob, 1, ndim), 1, fc, 0)
        if fused:
            assert mg.set_fused(True, min_cells=256) >= 1
        return mg
    together = solver(4).solve(rhs, ncycles=3, symmetrized=True)
    one = solver(1)
    separate = np.stack([one.solve(rhs[m], ncycles=3, symmetrized=True) for m in range(4)])
    assert together.shape == separate.shape and np.array_equal(together, separate)


@pytest.mark.parametrize("n,dx", [((64, 48), (0.3, 0.2)), ((32, 16, 24), (0.5, 0.4, 0.25))])
def test_zero_slope_boundaries_give_the_neumann_operator_and_the_vcycle_contracts(n, dx):
    """ampe_mg_set_zero_slope / HostMG::setZeroSlope (the blocks of a deck with boundary_N = "slope", "0"): the face
    coefficient of every boundary face is zero on every level -- the operator is the finite-volume operator with no
    flux through the boundary (checked against numpy), a constant is in the null space of its diffusion part, and
    the V-cycle contracts"""
    from oracle import pyoracle
    nd = len(n)
    shape = tuple(reversed(n))
    rng = np.random.default_rng(1)
    h = pyoracle.HostMG(n, dx)
    h.set_zero_slope([1] * nd)
    dcoef = -0.37
    h.set_elliptic(m_const=1.0, c_const=1.0, d_const=dcoef)
    u = rng.standard_normal(shape)
    ref = u.copy()
    for a in range(nd):
        ax = nd - 1 - a
        up = np.roll(u, -1, ax) - u
        dn = u - np.roll(u, 1, ax)
        last = [slice(None)] * nd
        last[ax] = -1
        first = [slice(None)] * nd
        first[ax] = 0
        up[tuple(last)] = 0.0
        dn[tuple(first)] = 0.0
        ref += dcoef / dx[a] ** 2 * (up - dn)
    got = h.apply(u)
    assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max()
    one = np.ones(shape)
    assert np.abs(h.apply(one) - one).max() <= 1e-13          # C u only: no flux anywhere
    rhs = rng.standard_normal(shape)
    res = []
    for nc in (1, 2, 3):
        z = h.solve(rhs, ncycles=nc)
        res.append(np.linalg.norm(rhs - h.apply(z)) / np.linalg.norm(rhs))
    assert res[0] < 0.15 and res[1] < 0.3 * res[0] + 1e-12 and res[2] < 0.3 * res[1] + 1e-12
    # one periodic direction, one zero-slope direction: different operators
    hp = pyoracle.HostMG(n, dx)
    hp.set_elliptic(m_const=1.0, c_const=1.0, d_const=dcoef)
    assert np.abs(hp.apply(u) - got).max() > 1e-3
