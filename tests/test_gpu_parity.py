"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI of
libampe_b200.so, against the CPU oracle on the same seeded inputs, against the
committed golden vectors, and -- at BASELINE.json's full sizes -- through
size-independent properties.

Tolerance: north_star's per-cell relative error 1e-12 in fp64, measured as
|gpu-ref| / max(|ref|, 1e-3 ||ref||_inf) (parity.rel_err).  The ill-conditioned outputs (composition RHS of
the CALPHAD configurations, symmetry-aware quaternion RHS) are judged by the extended-precision arbiter
(parity.check): the device must be as close to the long-double evaluation of the reference's formulas as
the fp64 restatement itself is."""
import numpy as np
import pytest
import torch

import parity
from test_oracle_golden import load_golden

pytestmark = pytest.mark.gpu

TOL = parity.TOL


def _check(errs, cfg=None):
    parity.check(errs)


@pytest.mark.parametrize("name", list(parity.SMALL))
def test_rhs_matches_oracle(name):
    cfg, st = parity.make_case(name)
    rot = parity.random_rotations(cfg) if cfg.symmetry_aware else None
    errs = parity.compare(name, cfg, st, fd_flags=(0, 1, 0), rotations=rot)
    _check(errs, cfg)


@pytest.mark.parametrize("name", list(parity.SMALL))
def test_rhs_matches_golden(name):
    cfg, st, rot, g = load_golden(name)
    outs, extra, launches = parity.run_gpu(cfg, st, fd_flags=(0,), rotations=rot)
    assert launches >= 1
    errs = {}
    for k in ("phase", "quat", "conc", "temperature"):
        if ("ydot_" + k) in g:
            errs["fd0:" + k] = parity.rel_err(outs[0][k], g["ydot_" + k])
    if extra is not None:
        errs["cl"] = parity.rel_err(extra[0], g["cl"])
        errs["ca"] = parity.rel_err(extra[1], g["ca"])
    errs = parity.Errs(errs)
    if parity.needs_arbiter(cfg):
        ld, ld_extra = parity.run_arbiter(cfg, st, (0,), rot)
        for k in ("phase", "quat", "conc", "temperature"):
            if ("ydot_" + k) in g:
                errs.ld["fd0:" + k] = (parity.rel_err_ld(outs[0][k], ld[0][k]), parity.rel_err_ld(g["ydot_" + k], ld[0][k]))
        if extra is not None:
            errs.ld["cl"] = (parity.rel_err_ld(extra[0], ld_extra[0]), parity.rel_err_ld(g["cl"], ld_extra[0]))
            errs.ld["ca"] = (parity.rel_err_ld(extra[1], ld_extra[1]), parity.rel_err_ld(g["ca"], ld_extra[1]))
    _check(errs, cfg)
    if "ydot_conc" in g and cfg.free_energy == 2:
        ref = g["ydot_conc"]
        assert np.abs(outs[0]["conc"] - ref).max() / np.abs(ref).max() < 1e-12


@pytest.mark.parametrize("name,kw", [
    ("dendrite2d", dict(nx=40, ny=33)),      # not multiples of the 32x16 tile
    ("dendrite2d", dict(nx=31, ny=17)),
    ("auni2d", dict(nx=33, ny=47)),
    ("gg3d_hbsm", dict(nx=33, ny=9, nz=7)),  # ragged 3D tiles (32x4x4)
    ("auni3d", dict(nx=20, ny=6, nz=5)),
    ("pfhub1a", dict(nx=16, ny=12)),
])
def test_ragged_sizes(name, kw):
    cfg, st = parity.make_case(name, **kw)
    rot = parity.random_rotations(cfg) if cfg.symmetry_aware else None
    errs = parity.compare(name, cfg, st, fd_flags=(0,), rotations=rot)
    _check(errs, cfg)


@pytest.mark.parametrize("variant", ["no_symmetry", "modulus_from_sides", "floor_tanh", "floor_sqrt",
                                     "harmonic_avg", "lag_off", "isotropic_flux", "frozen_quat",
                                     "anisotropic_3d_kks", "anisotropic_3d_ebs", "anisotropic_3d_clamped",
                                     "kks_flux_calphad", "kks_flux_calphad_3d"])
def test_model_switches(variant):
    """the runtime switches of QuatModelParameters that the hot path honours"""
    name = "auni2d"
    if variant in ("isotropic_flux", "frozen_quat"):
        name = "dendrite2d"
    if variant in ("anisotropic_3d_kks", "anisotropic_3d_clamped"):
        name = "gg3d_hbsm"
    if variant == "anisotropic_3d_ebs":
        name = "auni3d"
    if variant == "kks_flux_calphad_3d":
        name = "auni3d"
    cfg, st = parity.make_case(name)
    rot = None
    if variant.startswith("kks_flux_calphad"):
        # rhs_form "kks" with the CALPHAD free energy (tests/KKScomposition): Arrhenius D_solid / D_liquid instead of the
        # CALPHAD mobilities, the KKS kernel's Newton and driving force as for "ebs"
        cfg.symmetry_aware = 0
        cfg.conc_rhs_form = 2
        cfg.D_solid, cfg.D_liquid = 0.125, 1224.23
        cfg.Q0_solid, cfg.Q0_liquid = 1.0e4, 2.0e4
    if variant.startswith("anisotropic_3d"):
        # 3D anisotropic interface energy (3d/quatrhs.m4:149-349) inside the fused marching kernel
        cfg.symmetry_aware = 0
        cfg.phase_flux_type = 2
        cfg.epsilon_anisotropy = 0.05
        if variant == "anisotropic_3d_clamped":
            cfg.zero_slope[0] = cfg.zero_slope[2] = 1
    if variant == "no_symmetry":
        cfg.symmetry_aware = 0
    elif variant == "modulus_from_sides":
        cfg.symmetry_aware = 0
        cfg.quat_grad_modulus_from_cells = 0
    elif variant == "floor_tanh":
        cfg.symmetry_aware = 0
        cfg.grad_floor_type = b"t"
    elif variant == "floor_sqrt":
        cfg.symmetry_aware = 0
        cfg.grad_floor_type = b"s"
    elif variant == "harmonic_avg":
        cfg.symmetry_aware = 0
        cfg.avg_func = b"h"
        cfg.conc_avg_func = b"h"
    elif variant == "lag_off":
        cfg.symmetry_aware = 0
        cfg.lag_quat_sidegrad = 0
    elif variant == "isotropic_flux":
        cfg.phase_flux_type = 1
    elif variant == "frozen_quat":
        cfg.evolve_quat = 0   # H_parameter == 0: orientation only feeds the anisotropy
    if cfg.symmetry_aware:
        rot = parity.random_rotations(cfg)
    errs = parity.compare(name, cfg, st, fd_flags=(0, 1), rotations=rot)
    _check(errs, cfg)


def test_fd_flag_lagging_semantics_gpu():
    """fd_flag=1 on a perturbed state reuses the lagged 1/|grad q| and the lagged
    composition diffusivities (QuatIntegrator.cc:3183-3189, 3268-3269)"""
    from ampe_b200 import rhs
    from oracle import pyoracle
    name = "auni2d"
    cfg, st = parity.make_case(name)
    cfg.symmetry_aware = 0
    y = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    st2 = {k: (None if v is None else v.clone()) for k, v in st.items()}
    st2["quat"] = parity.fields.smooth_unit(st["quat"] + 1e-3 * torch.roll(st["quat"], 3, -1))
    st2["phase"] = (st["phase"] + 1e-4 * torch.sin(torch.arange(st["phase"].numel(), dtype=torch.float64)).reshape(st["phase"].shape)).contiguous()
    y2 = {k: (None if v is None else v.numpy().copy()) for k, v in st2.items()}
    o = pyoracle.Oracle(cfg)
    o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
    o.eval(0.0, y, 0)
    _, o_lag = o.eval(0.0, y2, 1)
    _, o_full = o.eval(0.0, y2, 0)
    r = rhs.QuatIntegratorRHS(cfg)
    yg, yg2 = rhs.to_device(st), rhs.to_device(st2)
    c0 = yg["conc"].reshape(-1).clone()
    r.resetRefPhaseConcentrations(c0, c0.clone())
    out = yg.like()
    r.evaluateRHSFunction(0.0, yg, out, 0)
    lag = yg.like()
    r.evaluateRHSFunction(0.0, yg2, lag, 1)
    full = yg.like()
    r.evaluateRHSFunction(0.0, yg2, full, 0)
    torch.cuda.synchronize()
    # the same call sequence in long double: arbiter for the composition RHS
    a = pyoracle.OracleLD(cfg)
    a.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
    a.eval(0.0, y, 0)
    _, a_lag = a.eval(0.0, y2, 1)
    _, a_full = a.eval(0.0, y2, 0)
    a.close()
    for k in ("phase", "quat", "conc"):
        parity.check_one(k + ":lagged", lag[k].cpu().numpy(), o_lag[k], a_lag[k])
        parity.check_one(k + ":full", full[k].cpu().numpy(), o_full[k], a_full[k])
    assert not np.array_equal(o_lag["quat"], o_full["quat"])
    assert not torch.equal(lag["quat"], full["quat"])
    assert not torch.equal(lag["conc"], full["conc"])
    # y must not be modified (QuatIntegrator.h:202)
    assert torch.equal(yg2["phase"].cpu(), st2["phase"])


def test_error_paths():
    from ampe_b200 import configs, rhs
    from ampe_b200.lib import AmpeError
    cfg = configs.dendrite2d(nx=64, ny=64)
    cfg.qlen = 3
    with pytest.raises(AmpeError):
        rhs.QuatIntegratorRHS(cfg)
    cfg = configs.auni2d(nx=64, ny=64)
    r = rhs.QuatIntegratorRHS(cfg)
    st = parity.fields.make_state("auni2d", cfg)
    y = rhs.to_device(st)
    with pytest.raises(AmpeError):   # Newton reference not set
        r.evaluateRHSFunction(0.0, y, y.like(), 0)
    cfg = configs.dendrite2d(nx=64, ny=64)
    r = rhs.QuatIntegratorRHS(cfg)
    y = rhs.to_device(parity.fields.make_state("dendrite2d", cfg))
    with pytest.raises(AmpeError):   # fd_flag=1 before any full evaluation
        r.evaluateRHSFunction(0.0, y, y.like(), 1)


def test_host_buffer_entry_point():
    """ampe_rhs_eval_host: host y -> device, evaluate, ydot -> host"""
    from ampe_b200 import rhs
    cfg, st = parity.make_case("dendrite2d")
    o_outs, _ = parity.run_oracle(cfg, st)
    r = rhs.QuatIntegratorRHS(cfg)
    yh = {k: (None if v is None else v.clone().pin_memory()) for k, v in st.items()}
    ydh = {k: (None if v is None else torch.zeros_like(v).pin_memory()) for k, v in st.items()}
    r.evaluateRHSFunctionHost(0.0, yh, ydh, 0)
    for k in ("phase", "quat", "temperature"):
        assert parity.rel_err(ydh[k].numpy(), o_outs[0][1][k]) <= TOL


@pytest.mark.parametrize("name", list(parity.SMALL))
@pytest.mark.parametrize("chunks", [1, 3])
def test_pipelined_host_entry_point_equals_device_path(name, chunks, monkeypatch):
    """ampe_rhs_eval_host streams the slab in chunks (H2D | kernels | D2H overlapped): bit-identical
    to the device-resident evaluation, for full and lagged (fd_flag=1) evaluations"""
    from ampe_b200 import rhs
    monkeypatch.setenv("AMPE_B200_HOST_CHUNKS", str(chunks))
    cfg, st = parity.make_case(name)
    rot = parity.random_rotations(cfg) if cfg.symmetry_aware else None
    y = rhs.to_device(st)
    rd, rh = rhs.QuatIntegratorRHS(cfg), rhs.QuatIntegratorRHS(cfg)
    for r in (rd, rh):
        if cfg.conc_rhs_form in (2, 3):
            c0 = y["conc"].reshape(-1).clone()
            r.resetRefPhaseConcentrations(c0, c0.clone())
        if cfg.symmetry_aware:
            r.setSymmetryRotations([torch.as_tensor(a).cuda() for a in rot])
    yh = {k: (None if v is None else v.clone().pin_memory()) for k, v in st.items()}
    for fd in (0, 1, 0):
        ydh = {k: (None if v is None else torch.full_like(v, float("nan")).pin_memory()) for k, v in st.items()}
        ref = y.like()
        rd.evaluateRHSFunction(0.0, y, ref, fd)
        rh.evaluateRHSFunctionHost(0.0, yh, ydh, fd)
        torch.cuda.synchronize()
        for k, v in ref.items():
            if v is not None and not (k == "quat" and not cfg.evolve_quat):
                assert torch.equal(v.cpu().reshape(-1), ydh[k].reshape(-1)), (name, fd, k)
    if cfg.conc_rhs_form in (2, 3):
        assert rh.newtonFailures() == 0


@pytest.mark.parametrize("name,kw", [("dendrite2d", dict(nx=256, ny=200)), ("auni2d", dict(nx=192, ny=136)),
                                     ("dendrite2d", dict(nx=130, ny=97))])
def test_tma_staging_equals_cp_async_staging(name, kw, monkeypatch):
    """the opt-in persistent TMA kernel (AMPE_B200_TMA=1: interior tiles by cp.async.bulk.tensor,
    boundary tiles by cp.async) and the default tile kernel produce the same bits: full, lagged and
    split launches"""
    from ampe_b200 import rhs
    cfg, st = parity.make_case(name, **kw)
    cfg.symmetry_aware = 0
    y = rhs.to_device(st)
    outs = {}
    for mode in ("tma", "cp_async"):
        if mode == "tma":
            monkeypatch.setenv("AMPE_B200_TMA", "1")
        else:
            monkeypatch.delenv("AMPE_B200_TMA", raising=False)
        r = rhs.QuatIntegratorRHS(cfg)
        if cfg.conc_rhs_form in (2, 3):
            c0 = y["conc"].reshape(-1).clone()
            r.resetRefPhaseConcentrations(c0, c0.clone())
        res = []
        for fd, parts in ((0, (0,)), (1, (0,)), (0, (1, 2))):
            yd = y.like()
            for part in parts:
                r.evaluateRHSFunction(0.0, y, yd, fd, part=part)
            res.append(yd)
        torch.cuda.synchronize()
        outs[mode] = res
        r.close()
    for a, b in zip(outs["tma"], outs["cp_async"]):
        for k, v in a.items():
            if v is not None and not (k == "quat" and not cfg.evolve_quat):
                assert torch.isfinite(v).all(), k
                assert torch.equal(v, b[k]), (name, k)


def test_split_evaluation_equals_full():
    """interior + boundary launches (used to overlap the halo exchange) == one full launch"""
    from ampe_b200 import rhs
    for name in ("dendrite2d", "auni3d", "gg3d_hbsm", "pfhub1a"):
        cfg, st = parity.make_case(name)
        y = rhs.to_device(st)
        r = rhs.QuatIntegratorRHS(cfg)
        if cfg.conc_rhs_form in (2, 3):
            c0 = y["conc"].reshape(-1).clone()
            r.resetRefPhaseConcentrations(c0, c0.clone())
        full, split = y.like(), y.like()
        r.evaluateRHSFunction(0.0, y, full, 0)
        r.evaluateRHSFunction(0.0, y, split, 0, part=1)
        r.evaluateRHSFunction(0.0, y, split, 0, part=2)
        torch.cuda.synchronize()
        for k, v in full.items():
            if v is not None and not (k == "quat" and not cfg.evolve_quat):
                assert torch.equal(v, split[k]), (name, k)


FULL = {
    "dendrite2d": dict(nx=2048, ny=2048),
    "auni2d": dict(nx=4096, ny=4096),
    "gg3d_hbsm": dict(nx=512, ny=512, nz=128),
    "auni3d": dict(nx=512, ny=256, nz=128),
}


@pytest.mark.parametrize("name", list(FULL))
def test_full_size_properties(name):
    """BASELINE.json sizes (3D: one GPU's share): projection q.ydot_q = 0, conservation
    sum ydot_c = 0, determinism, and periodic translation invariance by a shift that is not
    a multiple of the tile (exercises every wrap / tile-edge path at scale)."""
    from ampe_b200 import configs, rhs
    cfg = configs.BUILDERS[name](**FULL[name])
    cfg.symmetry_aware = 0
    st = parity.fields.make_state(name, cfg, device="cuda")
    y = rhs.SolutionVector(st)
    r = rhs.QuatIntegratorRHS(cfg)
    if cfg.conc_rhs_form in (2, 3):
        c0 = y["conc"].reshape(-1).clone()
        r.resetRefPhaseConcentrations(c0, c0.clone())
    out = y.like()
    r.evaluateRHSFunction(0.0, y, out, 0)
    again = y.like()
    r.evaluateRHSFunction(0.0, y, again, 0)
    torch.cuda.synchronize()
    assert r.newtonFailures() == 0
    for k, v in out.items():
        if v is not None:
            assert torch.isfinite(v).all(), k
            assert torch.equal(v, again[k]), k            # idempotent / deterministic
    q, yq = y["quat"], out["quat"]
    dot = (q * yq).sum(0).abs().max().item()
    assert dot / (yq.abs().max().item() + 1e-300) < 1e-9
    if out["conc"] is not None:
        s = out["conc"].sum().abs().item() / out["conc"].abs().sum().item()
        assert s < 1e-9
    # translation invariance: RHS(shift(y)) == shift(RHS(y)) bit for bit
    shifts = (5, 3, 7)
    dims = (-1, -2, -3)[:cfg.ndim]
    ys = rhs.SolutionVector({k: (None if v is None else torch.roll(v, shifts[:cfg.ndim], dims).contiguous())
                             for k, v in y.items()})
    if cfg.conc_rhs_form in (2, 3):
        c0 = ys["conc"].reshape(-1).clone()
        r.resetRefPhaseConcentrations(c0, c0.clone())
    outs = ys.like()
    r.evaluateRHSFunction(0.0, ys, outs, 0)
    torch.cuda.synchronize()
    for k, v in out.items():
        if v is not None:
            assert torch.equal(torch.roll(v, shifts[:cfg.ndim], dims), outs[k]), k


@pytest.mark.parametrize("name", ["dendrite2d", "auni2d", "gg3d_hbsm", "auni3d", "pfhub1a", "gg3d_hbsm:anisotropic"])
def test_slab_decomposition_equals_single_rank(name):
    """two slab 'ranks' on one GPU, ghost planes handed over through ampe_rhs_set_halo (what
    halo.py receives from the neighbour over NCCL): the union of the two slab evaluations is
    bit-identical to the single-rank evaluation of the whole periodic domain, for the full
    launch and for the interior/boundary split used to overlap the exchange"""
    from ampe_b200 import configs, rhs
    from ampe_b200.halo import slab_dim, slab_planes
    name, _, flux = name.partition(":")
    cfg, st = parity.make_case(name)
    cfg.symmetry_aware = 0
    if flux:
        cfg.phase_flux_type, cfg.epsilon_anisotropy = 2, 0.05
    ndim = cfg.ndim
    y = rhs.to_device(st)
    r = rhs.QuatIntegratorRHS(cfg)
    kks = cfg.conc_rhs_form in (2, 3)
    if kks:
        c0 = y["conc"].reshape(-1).clone()
        r.resetRefPhaseConcentrations(c0, c0.clone())
    ref = y.like()
    r.evaluateRHSFunction(0.0, y, ref, 0)
    ref1 = y.like()
    r.evaluateRHSFunction(0.0, y, ref1, 1)
    torch.cuda.synchronize()
    ns = cfg.n[ndim - 1]
    half = ns // 2
    ng = r.nghosts()
    dim = slab_dim(ndim)
    for rank in (0, 1):
        lo_i, hi_i = rank * half, (rank + 1) * half if rank == 0 else ns
        kw = dict(nx=cfg.n[0], ny=cfg.n[1])
        if ndim == 3:
            kw["nz"] = hi_i - lo_i
        else:
            kw["ny"] = hi_i - lo_i
        c2 = configs.BUILDERS[name](**kw)
        for d in range(3):
            c2.dx[d] = cfg.dx[d]
        c2.symmetry_aware = 0
        c2.phase_flux_type, c2.epsilon_anisotropy = cfg.phase_flux_type, cfg.epsilon_anisotropy
        c2.nranks, c2.rank = 2, rank
        take = lambda t, idx: t.index_select(dim, torch.tensor([i % ns for i in idx], device=t.device)).contiguous()
        ys = rhs.SolutionVector({k: (None if v is None else slab_planes(v, ndim, slice(lo_i, hi_i)).contiguous())
                                 for k, v in y.items()})
        lo = rhs.SolutionVector({k: (None if v is None else take(v, range(lo_i - ng, lo_i))) for k, v in y.items()})
        hi = rhs.SolutionVector({k: (None if v is None else take(v, range(hi_i, hi_i + ng))) for k, v in y.items()})
        r2 = rhs.QuatIntegratorRHS(c2)
        r2.setHalo(lo, hi)
        if kks:
            g = take(y["conc"], range(lo_i - ng, hi_i + ng))
            r2.setRefPhaseConcentrationsGhosted(g, g.clone())
        for split in (False, True):
            out = ys.like()
            if split:
                r2.evaluateRHSFunction(0.0, ys, out, 0, part=1)
                r2.evaluateRHSFunction(0.0, ys, out, 0, part=2)
            else:
                r2.evaluateRHSFunction(0.0, ys, out, 0)
            out1 = ys.like()
            r2.evaluateRHSFunction(0.0, ys, out1, 1)
            torch.cuda.synchronize()
            assert r2.newtonFailures() == 0
            for k, v in ref.items():
                if v is None or (k == "quat" and not cfg.evolve_quat):
                    continue
                assert torch.equal(out[k], slab_planes(v, ndim, slice(lo_i, hi_i))), (name, rank, split, k)
                assert torch.equal(out1[k], slab_planes(ref1[k], ndim, slice(lo_i, hi_i))), (name, rank, split, k, "fd1")
        r2.close()
    r.close()


@pytest.mark.parametrize("case", ["dendrite_deck", "single_grain_deck", "auni3d_all_clamped", "gg3d_mixed",
                                  "dendrite2d_x_only", "auni2d_ramp"])
def test_zero_slope_boundaries_and_temperature_ramp(case):
    """physical boundaries boundary_N = "slope", "0" (ghost = adjacent interior cell, corners included) per
    direction, the DeltaTemperature free energy and the ScalarTemperatureStrategy ramp T(t): the fused path against
    the restatement on the same fields, fd_flag 0 / 1 / 0, evaluated at a time t > 0"""
    from ampe_b200 import configs, fields, rhs
    from oracle import pyoracle
    time = 0.0
    if case == "dendrite_deck":
        cfg = configs.dendrite_test2d()
        base = "dendrite2d"
    elif case == "single_grain_deck":
        cfg = configs.single_grain_auni_test2d()
        base, time = "auni2d", 0.17
    elif case == "auni3d_all_clamped":
        cfg = configs.auni3d(nx=36, ny=20, nz=12)
        cfg.zero_slope[0] = cfg.zero_slope[1] = cfg.zero_slope[2] = 1
        base = "auni3d"
    elif case == "gg3d_mixed":
        cfg = configs.gg3d_hbsm(nx=36, ny=20, nz=12)
        cfg.zero_slope[1] = cfg.zero_slope[2] = 1       # periodic in x only
        cfg.dtemperaturedt, cfg.target_temperature, base, time = -20.0, 573.0, "gg3d_hbsm", 0.05
    elif case == "dendrite2d_x_only":
        cfg = configs.dendrite2d(nx=72, ny=56)
        cfg.zero_slope[0] = 1
        base = "dendrite2d"
    else:
        cfg = configs.auni2d(nx=72, ny=56, symmetry=False)
        cfg.dtemperaturedt, cfg.target_temperature, base, time = -200.0, 1220.0, "auni2d", 0.4
    # fields of the same family on this grid (several grains, smooth noise); the deck models reuse them
    c2 = configs.BUILDERS[base](**({"nx": cfg.n[0], "ny": cfg.n[1]} if cfg.ndim == 2 else
                                   {"nx": cfg.n[0], "ny": cfg.n[1], "nz": cfg.n[2]}))
    _, st = parity.make_case(base, **({"nx": cfg.n[0], "ny": cfg.n[1]} if cfg.ndim == 2 else
                                      {"nx": cfg.n[0], "ny": cfg.n[1], "nz": cfg.n[2]}))
    if cfg.qlen == 0:
        st["quat"] = None
    if not cfg.with_concentration:
        st["conc"] = None
    if cfg.with_unsteady_temperature and st.get("temperature") is None:
        st["temperature"] = (0.7 + 0.3 * st["phase"]).contiguous()
    if not cfg.with_unsteady_temperature:
        st["temperature"] = None
    y_np = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    fds = (0, 1, 0)
    outs = {}
    for who in ("oracle", "ld"):
        o = pyoracle.Oracle(cfg) if who == "oracle" else pyoracle.OracleLD(cfg)
        if cfg.conc_rhs_form in (2, 3):
            o.set_ref(y_np["conc"].ravel().copy(), y_np["conc"].ravel().copy())
        res = []
        for fd in fds:
            if who == "oracle":
                status, yd = o.eval(time, y_np, fd)
            else:
                status, yd = o.eval(time, y_np, fd)
            assert status == 0
            res.append(yd)
        outs[who] = (res, o.phase_concentrations() if cfg.conc_rhs_form in (2, 3) else None)
        o.close()
    y = rhs.to_device(st)
    r = rhs.QuatIntegratorRHS(cfg)
    if cfg.conc_rhs_form in (2, 3):
        c0 = y["conc"].reshape(-1).clone()
        r.resetRefPhaseConcentrations(c0, c0.clone())
    g = []
    for fd in fds:
        yd = y.like()
        r.evaluateRHSFunction(time, y, yd, fd)
        torch.cuda.synchronize()
        g.append({k: (None if v is None else v.cpu().numpy()) for k, v in yd.items()})
    gx = None
    if cfg.conc_rhs_form in (2, 3):
        cl, ca = r.phaseConcentrations()
        gx = (cl.cpu().numpy(), ca.cpu().numpy())
        assert r.newtonFailures() == 0
    r.close()
    errs = parity.compare_outputs(cfg, fds, outs["oracle"][0], outs["oracle"][1], g, gx,
                                  (outs["ld"][0], outs["ld"][1]))
    parity.check(errs)
    # the boundary really is not periodic: the result differs from the periodic model's
    if case in ("dendrite_deck", "auni3d_all_clamped"):
        cfgp = type(cfg).from_buffer_copy(cfg)
        cfgp.zero_slope[0] = cfgp.zero_slope[1] = cfgp.zero_slope[2] = 0
        rp = rhs.QuatIntegratorRHS(cfgp)
        if cfgp.conc_rhs_form in (2, 3):
            c0 = y["conc"].reshape(-1).clone()
            rp.resetRefPhaseConcentrations(c0, c0.clone())
        ydp = y.like()
        rp.evaluateRHSFunction(time, y, ydp, 0)
        torch.cuda.synchronize()
        assert not np.array_equal(ydp["phase"].cpu().numpy(), g[0]["phase"])
        rp.close()


@pytest.mark.parametrize("ndim", [2, 3])
def test_ebs_flux_with_quadratic_free_energy(ndim):
    """rhs_form "ebs" with the quadratic free energy and diffusion_type "temperature_dependent"
    (TbasedCompositionDiffusionStrategy: tests/OneGrainQuadratic, tests/ConservedVolume): Arrhenius diffusivity of each
    phase weighted with the phase fraction at the face, closed-form phase concentrations, quadratic driving force --
    against the restatement, fd_flag 0 / 1 / 0, at a time where the temperature ramp has moved T"""
    from ampe_b200 import configs
    cfg = configs.one_grain_quadratic_test(ndim)
    base = "auni2d" if ndim == 2 else "gg3d_hbsm"
    small = parity.SMALL[base]
    cfg.n[0], cfg.n[1] = small["nx"], small["ny"]
    if ndim == 3:
        cfg.n[2] = small["nz"]
    _, st0 = parity.make_case("gg3d_hbsm" if ndim == 3 else "auni2d")
    phi = st0["phase"]
    # compositions in the range of the deck (0.06 liquid, 0.1 solid)
    from ampe_b200 import fields
    h = fields.h_pbg(phi)
    conc = (0.1 * h + 0.06 * (1 - h) + fields.smooth_noise(tuple(phi.shape), 1e-3, "cpu", 5)).contiguous()
    st = {"phase": phi, "quat": None, "conc": conc, "temperature": None}
    errs = parity.compare(base, cfg, st, fd_flags=(0, 1, 0))
    _check(errs, cfg)
