"""SURVEY.md 8f rank 3 on the device: the block preconditioners behind include/ampe_b200_precond.h
(csrc/mg.cu) and the host mirror's CVSpgmrPrecondSet / CVSpgmrPrecondSolve.

* ampe_mg_apply against the reference's operators restated on the CPU (efo_compfluxvardc + efo_compresvarsca,
  set_j_ij + set_stencil): 1e-13.
* the device V-cycles against the host loop over the same per-cell functions (oracle/precond.cc part 2): the
  arithmetic is identical (--fmad=false / -ffp-contract=off, red-black order independent) -- 1e-12 -- and the
  residual measured by the RESTATED operator contracts.
* CVSpgmrPrecondSet on the five configurations: every level-0 coefficient array of every block equals the CPU
  context's (1e-12), CVSpgmrPrecondSolve equals the CPU solve (1e-10).
* the right-preconditioned implicit trajectory on the device against the same template driven by the CPU
  oracle: 1e-8, and fewer Krylov vectors than the unpreconditioned run.
Added at the end of round 1 without a GPU left; first executed (green) by the first GPU call of round 2."""
import os

import numpy as np
import pytest
import torch

import parity

# first executed on a B200 in round 2 (profiles/r02a_pytest_experiments.log): 32 green, part of the default GPU suite
pytestmark = pytest.mark.gpu

from test_oracle_precond import BLOCKS, _evolved, _ghosted, _random_elliptic, _side_from_lower  # noqa: E402


def _cuda(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n,dx", [((64, 48), (0.3, 0.2)), ((32, 16, 24), (0.5, 0.4, 0.25)), ((15, 9), (1.0, 1.0))])
def test_scalar_block_operator_and_solve(n, dx):
    from ampe_b200.precond import LevelSolver
    from oracle import pyoracle
    ndim = len(n)
    shape, m, c, lows, d = _random_elliptic(n, 3)
    d = [40.0 * x for x in d]
    rng = np.random.default_rng(4)
    u = rng.standard_normal(shape)
    mg_m, mg_c = _ghosted(m, 1, ndim), _ghosted(c, 2, ndim)
    g = LevelSolver(n, dx)
    g.set_elliptic(m=_cuda(mg_m), ngm=1, c=_cuda(mg_c), ngc=2, d=[_cuda(x) for x in d], ngd=0)
    ref = pyoracle.elliptic_apply(n, dx, mg_m, 1, mg_c, 2, d, u)
    got = g.apply(_cuda(u)).cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()
    h = pyoracle.HostMG(n, dx)
    h.set_elliptic(m=mg_m, ngm=1, c=mg_c, ngc=2, d=d, ngd=0)
    assert g.num_levels() == h.num_levels()
    for lvl in range(g.num_levels()):
        assert g.level_extents(lvl) == h.level_extents(lvl)
        for which in (0, 1, 3, 4) + ((5,) if ndim == 3 else ()):
            a, b = g.level_array(lvl, which).cpu().numpy(), h.level_array(lvl, which)
            assert np.abs(a - b).max() <= 1e-14 * np.abs(b).max(), (lvl, which)
    rhs = rng.standard_normal(shape)
    hist = []
    for nc in (1, 2, 3):
        z = g.solve(_cuda(rhs), ncycles=nc).cpu().numpy()
        zh = h.solve(rhs, ncycles=nc)
        assert np.abs(z - zh).max() <= 1e-12 * np.abs(zh).max(), nc
        hist.append(np.linalg.norm(rhs - pyoracle.elliptic_apply(n, dx, mg_m, 1, mg_c, 2, d, z)))
    assert hist[2] < hist[1] < hist[0] < np.linalg.norm(rhs)
    assert g.last_launch_count() > 0
    # constants, second diffusion array and scale (the EBS composition block)
    g.set_elliptic(m_const=0.7, c_const=1.0, d=[_cuda(x) for x in d], d2=[_cuda(x) for x in d], ngd=0, d_scale=-0.25)
    ones = np.ones(shape)
    ref2 = pyoracle.elliptic_apply(n, dx, 0.7 * ones, 0, ones, 0, [-0.5 * x for x in d], u)
    assert np.abs(g.apply(_cuda(u)).cpu().numpy() - ref2).max() <= 1e-13 * np.abs(ref2).max()
    g.close()


@pytest.mark.parametrize("n,dx", [((48, 40), (0.3, 0.2)), ((16, 12, 20), (0.5, 0.4, 0.25))])
def test_quaternion_block_operator_and_solve(n, dx):
    from ampe_b200.precond import LevelSolver
    from oracle import pyoracle
    ndim = len(n)
    rng = np.random.default_rng(5)
    shape = (n[2] if ndim == 3 else 1, n[1], n[0])
    mob = 0.1 + rng.random(shape)
    fc = [_side_from_lower(-(5.0 + 20.0 * rng.random(shape)), 2 - a) for a in range(ndim)]
    w = rng.standard_normal(shape)
    gamma = 0.37
    mob_g = _ghosted(mob, 1, ndim)
    g = LevelSolver(n, dx, with_column_scale=True)
    g.set_quat(gamma, _cuda(mob_g), 1, [_cuda(x) for x in fc], 0)
    ref = pyoracle.quat_stencil_apply(n, dx, gamma, np.sqrt(mob_g), 1, fc, w)
    assert np.abs(g.apply(_cuda(w)).cpu().numpy() - ref).max() <= 1e-13 * np.abs(ref).max()
    h = pyoracle.HostMG(n, dx, with_s=True)
    h.set_quat(gamma, mob_g, 1, fc, 0)
    rhs = rng.standard_normal(shape)
    z = g.solve(_cuda(rhs), ncycles=5, symmetrized=True).cpu().numpy()
    zh = h.solve(rhs, ncycles=5, symmetrized=True)
    assert np.abs(z - zh).max() <= 1e-12 * np.abs(zh).max()
    s = np.sqrt(mob)
    res = rhs / s - pyoracle.quat_stencil_apply(n, dx, gamma, np.sqrt(mob_g), 1, fc, z / s)
    # cell-to-cell random mobility and face coefficients: ~0.35 per cycle (0.005 after five on the CPU)
    assert np.linalg.norm(res) < 0.02 * np.linalg.norm(rhs / s)
    g.close()


def test_phasefacops_setc_kernel():
    from ampe_b200.precond import phasefacops_setc
    n = (20, 12, 8)
    rng = np.random.default_rng(6)
    phi = rng.random((8, 12, 20))
    m = 0.5 + rng.random((8, 12, 20))
    phi_g, m_g = _ghosted(phi, 2, 3), _ghosted(m, 1, 3)
    c = torch.zeros((8, 12, 20), dtype=torch.float64, device="cuda")
    phasefacops_setc(n, _cuda(phi_g), 2, _cuda(m_g), 1, 0.3, 1.7, "double", c, 0)
    expect = 1.0 + (0.3 * m) * 1.7 * (32.0 * (1.0 + 6.0 * phi * (phi - 1.0)))
    assert np.abs(c.cpu().numpy() - expect).max() <= 1e-14 * np.abs(expect).max()


def _device_context(name):
    from ampe_b200 import rhs
    from ampe_b200.host_rhs import HostQuatIntegrator
    cfg, st = parity.make_case(name)
    y = rhs.to_device(st)
    h = HostQuatIntegrator(cfg, True)
    if cfg.conc_rhs_form in (2, 3):
        c0 = y["conc"].reshape(-1).clone()
        h.resetRefPhaseConcentrations(c0, c0.clone())
    rot = parity.random_rotations(cfg) if cfg.symmetry_aware else None
    if rot is not None:
        h.setSymmetryRotations([torch.as_tensor(a).cuda() for a in rot])
    return cfg, st, y, h, rot


@pytest.mark.parametrize("name", ["dendrite2d", "auni2d", "gg3d_hbsm", "auni3d"])
def test_precond_set_and_solve_match_the_cpu_context(name):
    from oracle import pyoracle
    cfg, st, y, h, rot = _device_context(name)
    yd = y.like()
    h.evaluateRHSFunction(0.0, y, yd, 0)  # the fd_flag = 0 evaluation that precedes every set-up
    h.setupPreconditioners(2)
    gamma = 20 * parity.TRAJ_DT[name]
    h.CVSpgmrPrecondSet(0.0, y, gamma)
    torch.cuda.synchronize()
    yo = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    if cfg.conc_rhs_form in (2, 3):
        o.set_ref(yo["conc"].ravel().copy(), yo["conc"].ravel().copy())
    if rot is not None:
        o.set_rotations(rot)
    assert o.eval(0.0, yo, fd_flag=0)[0] == 0
    assert o.precond_setup(gamma, 2) == 0
    for k in _evolved(cfg):
        g, ho = h.preconditionerLevelSolver(BLOCKS[k]), o.precond_block(BLOCKS[k])
        assert g is not None and g.num_levels() == ho.num_levels()
        for which in (0, 1, 2, 3, 4, 5):
            b = ho.level_array(0, which)
            if b is None:
                continue
            a = g.level_array(0, which).cpu().numpy()
            # the composition diffusivities carry the Newton-solved c_l, c_a (1e-11 like the composition RHS)
            tol = 1e-10 if k == "conc" else 1e-12
            assert np.abs(a - b).max() <= tol * np.abs(b).max(), (k, which)
    rng = np.random.default_rng(31)
    r = {k: (None if v is None else rng.standard_normal(v.shape)) for k, v in yo.items()}
    rc, zo = o.precond_solve(r)
    assert rc == 0
    rd = y.like()
    zd = y.like()
    for k in ("phase", "quat", "conc", "temperature"):
        if r.get(k) is not None and rd[k] is not None:
            rd[k].copy_(torch.as_tensor(r[k]))
    h.CVSpgmrPrecondSolve(rd, zd)
    torch.cuda.synchronize()
    for k in _evolved(cfg):
        z = zd[k].cpu().numpy()
        assert np.abs(z - zo[k]).max() <= 1e-10 * np.abs(zo[k]).max(), k
    assert h.precondStats() == {"precond_setups": 1.0, "precond_solves": 1.0}
    o.close()
    h.close()


@pytest.mark.parametrize("name,mult,nsteps", [("dendrite2d", 50, 4), ("auni2d", 20, 3), ("gg3d_hbsm", 40, 3),
                                              ("auni3d", 20, 3)])
def test_preconditioned_trajectory_matches_oracle_backend(name, mult, nsteps):
    from oracle import pyoracle
    kw = dict(order=2, rtol=1e-8, atol=1e-10, max_krylov=30, max_newton=8)
    dt = parity.TRAJ_DT[name] * mult
    cfg, st, y, h, rot = _device_context(name)
    h.setupPreconditioners(2)
    rc, sg = h.integrateImplicit(y, dt, nsteps, **kw)
    torch.cuda.synchronize()
    sg.update(h.precondStats())
    yo = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    if cfg.conc_rhs_form in (2, 3):
        o.set_ref(yo["conc"].ravel().copy(), yo["conc"].ravel().copy())
    if rot is not None:
        o.set_rotations(rot)
    o.set_preconditioner(2)
    rco, so = o.integrate_implicit(yo, dt, nsteps, **kw)
    so.update(o.precond_stats())
    assert rc == 0 and rco == 0, (sg, so)
    assert sg["precond_setups"] == sg["newton_iterations"] > 0
    for k in ("phase", "quat", "conc", "temperature"):
        if yo.get(k) is None:
            continue
        scale = max(np.abs(yo[k]).max(), 1e-300)
        assert np.abs(y[k].cpu().numpy() - yo[k]).max() <= 1e-8 * scale, (k, sg, so)
    # unpreconditioned device run of the same steps: more Krylov vectors
    cfg, st, y2, h2, rot = _device_context(name)
    rc2, s2 = h2.integrateImplicit(y2, dt, nsteps, **kw)
    assert rc2 == 0 and sg["linear_iterations"] < s2["linear_iterations"], (sg, s2)
    o.close()
    h.close()
    h2.close()


@pytest.mark.parametrize("name", ["auni2d", "gg3d_hbsm", "auni3d"])
def test_dquatdphi_block_matches_the_cpu_restatement(name):
    """QuatFACOps::multiplyDQuatDPhiBlock through the piecewise kernels (QUATMOBILITYDERIV, QUATDIFFUSIONDERIV,
    COMPUTE_DQUATDPHI_FACE_COEF, COMPUTE_FLUX, ADD_QUAT_OP, MULTICOMPONENT_MULTIPLY) against the restatement,
    and the coupled CVSpgmrPrecondSolve against the CPU one"""
    from oracle import pyoracle
    cfg, st, y, h, rot = _device_context(name)
    yd = y.like()
    h.evaluateRHSFunction(0.0, y, yd, 0)
    h.setupPreconditioners(2, precond_has_dquatdphi=True)
    gamma = 20 * parity.TRAJ_DT[name]
    h.CVSpgmrPrecondSet(0.0, y, gamma)
    yo = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    if cfg.conc_rhs_form in (2, 3):
        o.set_ref(yo["conc"].ravel().copy(), yo["conc"].ravel().copy())
    if rot is not None:
        o.set_rotations(rot)
    assert o.eval(0.0, yo, fd_flag=0)[0] == 0
    assert o.precond_setup(gamma, 2, dquatdphi=True) == 0
    rng = np.random.default_rng(43)
    z = rng.standard_normal(yo["phase"].shape)
    ref = o.precond_dquatdphi(z)
    got = h.multiplyDQuatDPhiBlock(_cuda(z), cfg.qlen).cpu().numpy()
    assert np.abs(got.reshape(ref.shape) - ref).max() <= 1e-11 * np.abs(ref).max()
    r = {k: (None if v is None else rng.standard_normal(v.shape)) for k, v in yo.items()}
    rc, zo = o.precond_solve(r)
    rd, zd = y.like(), y.like()
    for k in ("phase", "quat", "conc", "temperature"):
        if r.get(k) is not None and rd[k] is not None:
            rd[k].copy_(torch.as_tensor(r[k]))
    h.CVSpgmrPrecondSolve(rd, zd)
    torch.cuda.synchronize()
    for k in _evolved(cfg):
        assert np.abs(zd[k].cpu().numpy() - zo[k]).max() <= 1e-10 * np.abs(zo[k]).max(), k
    o.close()
    h.close()


@pytest.mark.parametrize("ndim", [2, 3])
def test_reference_kat_facpoisson_on_the_device(ndim):
    """tests/testFACPoisson.cc on the device solver (periodic image of the Dirichlet problem, see
    test_oracle_precond.facpoisson_case): max |computed - exact| < 1e-2"""
    from ampe_b200.precond import LevelSolver
    from test_oracle_precond import facpoisson_case
    n, dx, exact, rhs, nc = facpoisson_case(ndim)
    g = LevelSolver(n, dx)
    g.set_elliptic(m_const=1.0, c_const=5.0, d_const=-1.0)
    z = g.solve(_cuda(rhs), ncycles=10).cpu().numpy()
    assert np.abs(z - exact).max() < 1.0e-2
    res = rhs - g.apply(_cuda(z)).cpu().numpy()
    assert np.linalg.norm(res) < (1e-8 if ndim == 2 else 1e-6) * np.linalg.norm(rhs)
    g.close()


def test_reference_kat_phasefac_on_the_device():
    """tests/testPhaseFAC.cc on the device: PhaseFACOps::setC kernel + solver, max |computed - exact| < 1e-2"""
    from ampe_b200.precond import LevelSolver, phasefacops_setc
    from test_oracle_precond import phasefac_case
    k = phasefac_case()
    ny = 8
    n, dx = (4 * k["nc"], ny), (k["h"], k["h"])
    tile = lambda a: np.ascontiguousarray(np.tile(a, (1, ny, 1)))
    exact, rhs, phi = tile(k["exact"]), tile(k["rhs"]), tile(k["phi_coef"])
    mob = _cuda(np.full(phi.shape, k["mob"]))
    c = torch.empty(phi.shape, dtype=torch.float64, device="cuda")
    phasefacops_setc(n, _cuda(phi), 0, mob, 0, k["gamma"], k["w"], "double", c, 0)
    g = LevelSolver(n, dx)
    g.set_elliptic(m=mob, ngm=0, c=c, ngc=0, d_const=-k["gamma"] * k["eps"] ** 2)
    z = g.solve(_cuda(rhs), ncycles=10).cpu().numpy()
    assert np.abs(z - exact)[0, :, :k["nc"]].max() < 1.0e-2
    g.close()


@pytest.mark.parametrize("n,dx", [((256, 192), (1.0, 1.0)), ((64, 32, 48), (1.0, 1.0, 1.0)), ((48, 40), (1.0, 0.5))])
def test_one_block_tail_and_graph_replay_are_bit_identical(n, dx):
    """the coarse levels inside one block (default) == one launch per phase (AMPE_B200_MG_TAIL=0), bit for bit,
    with fewer launches; the fused red-black tile pass (default) == the two colour half-sweeps
    (AMPE_B200_MG_FUSED=0); the captured solve (AMPE_B200_MG_GRAPH=1, explicit stream) replays to the same bits"""
    from ampe_b200.precond import LevelSolver
    shape, m, c, lows, d = _random_elliptic(n, 17)
    d = [30.0 * x for x in d]
    rhs = _cuda(np.random.default_rng(18).standard_normal(shape))
    outs, launches = {}, {}
    for mode, env in (("tail", {}), ("levels", {"AMPE_B200_MG_TAIL": "0"}), ("unfused", {"AMPE_B200_MG_FUSED": "0"}),
                      ("graph", {"AMPE_B200_MG_GRAPH": "1"})):
        old = {k: os.environ.get(k) for k in ("AMPE_B200_MG_TAIL", "AMPE_B200_MG_GRAPH", "AMPE_B200_MG_FUSED")}
        for k in old:
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            g = LevelSolver(n, dx)  # the switches are read when the solver is created
        finally:
            for k, v in old.items():
                os.environ.pop(k, None)
                if v is not None:
                    os.environ[k] = v
        g.set_elliptic(m=_cuda(m), ngm=0, c=_cuda(c), ngc=0, d=[_cuda(x) for x in d], ngd=0)
        stream = torch.cuda.Stream() if mode == "graph" else None
        if stream is not None:
            stream.wait_stream(torch.cuda.current_stream())
        z = g.solve(rhs, ncycles=3, stream=stream)
        if mode == "graph":
            stream.synchronize()
            z2 = torch.empty_like(z)
            z2.copy_(z)
            g.solve(rhs, ncycles=3, out=z, stream=stream)  # second call: replay
            stream.synchronize()
            assert torch.equal(z, z2)
        torch.cuda.synchronize()
        outs[mode], launches[mode] = z.clone(), g.last_launch_count()
        g.close()
    assert torch.equal(outs["tail"], outs["levels"])
    assert torch.equal(outs["tail"], outs["unfused"])
    assert torch.equal(outs["tail"], outs["graph"])
    assert launches["tail"] < launches["levels"]


@pytest.mark.parametrize("name,mult,precond", [("pfhub1a", 200, 0), ("dendrite2d", 300, 2), ("gg3d_hbsm", 120, 2)])
def test_adaptive_integration_matches_oracle_backend(name, mult, precond):
    """ImplicitIntegrator::advanceTo on device vectors (variable steps, local error test) against the same
    template driven by the CPU oracle: same number of steps and failures, fields within 1e-8"""
    from oracle import pyoracle
    kw = dict(rtol=1e-6, atol=1e-8, max_krylov=30, max_newton=4, max_steps=2000)
    dt = parity.TRAJ_DT[name]
    cfg, st, y, h, rot = _device_context(name)
    if precond:
        h.setupPreconditioners(precond)
    rc, sg = h.integrateAdaptive(y, mult * dt, dt, **kw)
    torch.cuda.synchronize()
    yo = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    if cfg.conc_rhs_form in (2, 3):
        o.set_ref(yo["conc"].ravel().copy(), yo["conc"].ravel().copy())
    if rot is not None:
        o.set_rotations(rot)
    o.set_preconditioner(precond)
    rco, so = o.integrate_adaptive(yo, mult * dt, dt, **kw)
    assert rc == 0 and rco == 0, (sg, so)
    assert sg["steps"] == so["steps"] and sg["error_test_failures"] == so["error_test_failures"], (sg, so)
    assert abs(sg["t_reached"] - mult * dt) <= 1e-12 * mult * dt
    for k in ("phase", "quat", "conc", "temperature"):
        if yo.get(k) is None:
            continue
        scale = max(np.abs(yo[k]).max(), 1e-300)
        assert np.abs(y[k].cpu().numpy() - yo[k]).max() <= 1e-8 * scale, (k, sg, so)
    o.close()
    h.close()


@pytest.mark.parametrize("name", ["dendrite2d", "auni3d", "pfhub1a"])
def test_scalar_diagnostics_match_the_restatement(name):
    """ampe_scalar_diagnostics (QuatModel::printScalarDiagnostics: solid fraction, integral / max concentration,
    Cex, temperature extrema / average, thermal energy) against the CPU restatement; the sums differ only by
    their order (1e-12), extrema are exact"""
    from ampe_b200 import rhs
    from oracle import pyoracle
    cfg, st = parity.make_case(name)
    y = rhs.to_device(st)
    r = rhs.QuatIntegratorRHS(cfg)
    got = r.printScalarDiagnostics(y)
    o = pyoracle.Oracle(cfg)
    ref = o.scalar_diagnostics({k: (None if v is None else v.numpy().copy()) for k, v in st.items()})
    o.close()
    assert set(got) == set(ref)
    for k in ref:
        if k in ("max_concentration", "min_temperature", "max_temperature"):
            assert got[k] == ref[k], k
        else:
            assert got[k] == pytest.approx(ref[k], rel=1e-12, abs=1e-13 * max(1.0, abs(ref[k]))), k
    r.close()


def test_cahnhilliard_regression_deck_on_the_device(tmp_path):
    """tests/CahnHilliard/test2d.py end to end on the device (see tests/test_regression_cahnhilliard.py):
    NetCDF initial conditions -> device, variable-step implicit integration to t = 300 with
    printScalarDiagnostics every 10 time units; consecutive composition integrals within 1e-5 relative"""
    from ampe_b200 import host_rhs, netcdf_classic, rhs
    from test_regression_cahnhilliard import deck, make_initial
    cfg = deck()
    path = str(tmp_path / "64x64.nc")
    netcdf_classic.write(path, {"concentration0": make_initial(64, 64)})
    y = rhs.to_device(host_rhs.read_initial_conditions(path, cfg))
    h = host_rhs.HostQuatIntegrator(cfg, True)
    diag = rhs.QuatIntegratorRHS(cfg)
    atol, old, t, step = 1.0e-4, -1.0, 0.0, 1.0e-3
    while t < 300.0:
        rc, st = h.integrateAdaptive(y, t + 10.0, step, t0=t, rtol=1e-2 * atol, atol=atol, max_steps=5000)
        assert rc == 0, st
        t, step = st["t_reached"], st["last_step"]
        conc = diag.printScalarDiagnostics(y)["integral_concentration"]
        old = conc if old < 0.0 else old
        assert abs(conc - old) <= 1.0e-5 * conc, (t, conc, old)
        old = conc
    assert t >= 300.0
    c = y["conc"].cpu().numpy()
    assert c.max() - c.min() > 0.3
    h.close()
    diag.close()


@pytest.mark.parametrize("n,dx", [((128, 96), (0.3, 0.2)), ((32, 16, 24), (1.0, 1.0, 1.0))])
def test_components_solved_together_on_the_device(n, dx):
    """ampe_mg_create_multi: the four quaternion components in one solver (every pass updates all of them) ==
    four separate device solves bit for bit, == the host loop to 1e-12"""
    from ampe_b200.precond import LevelSolver
    from oracle import pyoracle
    ndim = len(n)
    shape = (n[2] if ndim == 3 else 1, n[1], n[0])
    mob = 0.1 + np.random.default_rng(61).random(shape)
    fc = [_side_from_lower(-(5.0 + 20.0 * np.random.default_rng(62 + a).random(shape)), 2 - a) for a in range(ndim)]
    rhs = np.random.default_rng(63).standard_normal((4,) + shape)
    mob_g = _ghosted(mob, 1, ndim)

    def device(ncomp):
        g = LevelSolver(n, dx, with_column_scale=True, ncomp=ncomp)
        g.set_quat(0.37, _cuda(mob_g), 1, [_cuda(x) for x in fc], 0)
        return g
    g4, g1 = device(4), device(1)
    together = g4.solve(_cuda(rhs), ncycles=3, symmetrized=True)
    separate = torch.stack([g1.solve(_cuda(rhs[m]), ncycles=3, symmetrized=True) for m in range(4)])
    assert torch.equal(together, separate)
    h = pyoracle.HostMG(n, dx, with_s=True, ncomp=4)
    h.set_quat(0.37, mob_g, 1, fc, 0)
    zh = h.solve(rhs, ncycles=3, symmetrized=True)
    assert np.abs(together.cpu().numpy() - zh).max() <= 1e-12 * np.abs(zh).max()
    g4.close()
    g1.close()
