"""The host-side Strategy mirror (C++, ampe_b200/host/) driving the PIECEWISE kernels
(include/ampe_b200_kernels.h) in the reference's own call order must reproduce the oracle,
and must agree with the fused path."""
import numpy as np
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu


def _run(cfg, st, use_fused, fd_flags, rot):
    from ampe_b200 import rhs
    from ampe_b200.host_rhs import HostQuatIntegrator
    y = rhs.to_device(st)
    h = HostQuatIntegrator(cfg, use_fused)
    if cfg.conc_rhs_form in (2, 3):
        c0 = y["conc"].reshape(-1).clone()
        h.resetRefPhaseConcentrations(c0, c0.clone())
    if cfg.symmetry_aware:
        h.setSymmetryRotations([torch.as_tensor(a).cuda() for a in rot])
    outs = []
    for fd in fd_flags:
        yd = y.like()
        h.evaluateRHSFunction(0.0, y, yd, fd)
        outs.append({k: (None if v is None else v.cpu().numpy()) for k, v in yd.items()})
    h.close()
    return outs


@pytest.mark.parametrize("name", list(parity.SMALL))
def test_strategy_path_matches_oracle(name):
    cfg, st = parity.make_case(name)
    rot = parity.random_rotations(cfg) if cfg.symmetry_aware else None
    fds = (0, 1)
    o_outs, _ = parity.run_oracle(cfg, st, fds, rot)
    ld = parity.run_arbiter(cfg, st, fds, rot)[0] if parity.needs_arbiter(cfg) else None
    for use_fused in (False, True):
        g = _run(cfg, st, use_fused, fds, rot)
        for n, ((status, yo), yg) in enumerate(zip(o_outs, g)):
            for k in ("phase", "quat", "conc", "temperature"):
                if yo.get(k) is None or (k == "quat" and not cfg.evolve_quat):
                    continue
                parity.check_one("%s fused=%s fd%d:%s" % (name, use_fused, fds[n], k), yg[k], yo[k],
                                 None if ld is None else ld[n][k])


def test_anisotropic_3d_piecewise_kernel():
    """3D anisotropic_gradient_flux (3d/quatrhs.m4:149-349) exists as a piecewise kernel; its
    flux must reduce to the isotropic gamma^2 grad(phi) when eps4 = 0"""
    import ctypes as C
    from ampe_b200 import lib
    L = lib.load()
    n = (12, 10, 8)
    ng = 1
    g = torch.Generator().manual_seed(3)
    phase = torch.rand((n[2] + 2, n[1] + 2, n[0] + 2), generator=g, dtype=torch.float64).cuda()
    quat = torch.rand((4, n[2] + 2, n[1] + 2, n[0] + 2), generator=g, dtype=torch.float64).cuda()
    dx = (C.c_double * 3)(0.1, 0.11, 0.12)
    lo = (C.c_int * 3)(0, 0, 0)
    hi = (C.c_int * 3)(n[0] - 1, n[1] - 1, n[2] - 1)
    fl = [torch.zeros((n[2] + (a == 2), n[1] + (a == 1), n[0] + (a == 0)), dtype=torch.float64).cuda()
          for a in range(3)]
    fa = [torch.zeros_like(f) for f in fl]
    pf = (C.c_void_p * 3)(*[f.data_ptr() for f in fl])
    pa = (C.c_void_p * 3)(*[f.data_ptr() for f in fa])
    eps = 0.25
    L.ampe_k_gradient_flux.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.ampe_k_anisotropic_gradient_flux.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                                   C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                   C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    assert L.ampe_k_gradient_flux(3, lo, hi, dx, eps, phase.data_ptr(), ng, pf, 0, None) == 0
    assert L.ampe_k_anisotropic_gradient_flux(3, lo, hi, dx, eps, 0.0, 4, phase.data_ptr(), ng,
                                              quat.data_ptr(), ng, 4, pa, 0, None) == 0
    torch.cuda.synchronize()
    for a in range(3):
        assert torch.allclose(fl[a], fa[a], rtol=1e-13, atol=1e-13)


def _side_shapes(n):
    ndim = len(n)
    shp = []
    for a in range(ndim):
        ext = [n[d] + (1 if d == a else 0) for d in range(ndim)]
        shp.append(tuple(reversed(ext)))
    return shp


@pytest.mark.parametrize("n,eps4", [((12, 10, 8), 0.05), ((9, 11, 7), 0.02)])
def test_anisotropic_3d_piecewise_kernel_vs_oracle(n, eps4):
    """3D anisotropic_gradient_flux (3d/quatrhs.m4:149-349) with eps4 != 0 against the restatement
    (oracle/kernels.cc anisotropic_gradient_flux), 1e-12 of the largest flux"""
    import ctypes as C
    from ampe_b200 import lib
    from oracle import pyoracle
    L, O = lib.load(), pyoracle.lib()
    ng = 1
    rng = np.random.default_rng(11)
    gshape = (n[2] + 2, n[1] + 2, n[0] + 2)
    phase = rng.random(gshape)
    quat = rng.standard_normal((4,) + gshape)
    quat /= np.sqrt((quat * quat).sum(0, keepdims=True))
    dx = (C.c_double * 3)(0.1, 0.11, 0.12)
    lo = (C.c_int * 3)(0, 0, 0)
    hi = (C.c_int * 3)(n[0] - 1, n[1] - 1, n[2] - 1)
    shp = _side_shapes(n)
    ref = [np.zeros(s) for s in shp]
    pr = (C.c_void_p * 3)(*[a.ctypes.data for a in ref])
    O.oracle_k_anisotropic_gradient_flux.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                                     C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                     C.c_int, C.c_int, C.c_void_p, C.c_int]
    O.oracle_k_anisotropic_gradient_flux.restype = None
    O.oracle_k_anisotropic_gradient_flux(3, lo, hi, dx, 0.25, eps4, 4, phase.ctypes.data, ng, quat.ctypes.data, ng,
                                         4, pr, 0)
    dphase, dquat = torch.as_tensor(phase).cuda(), torch.as_tensor(quat).cuda()
    got = [torch.zeros(s, dtype=torch.float64).cuda() for s in shp]
    pg = (C.c_void_p * 3)(*[f.data_ptr() for f in got])
    L.ampe_k_anisotropic_gradient_flux.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                                   C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                   C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    assert L.ampe_k_anisotropic_gradient_flux(3, lo, hi, dx, 0.25, eps4, 4, dphase.data_ptr(), ng,
                                              dquat.data_ptr(), ng, 4, pg, 0, None) == 0
    torch.cuda.synchronize()
    for a in range(3):
        scale = np.abs(ref[a]).max()
        assert scale > 0
        # the anisotropic term must matter, otherwise the comparison says nothing about eps4
        assert np.abs(got[a].cpu().numpy() - ref[a]).max() <= 1e-12 * scale, a


@pytest.mark.parametrize("ndim,three_phase", [(2, 0), (2, 1), (3, 1)])
def test_computerhspbg_full_argument_list(ndim, three_phase):
    """COMPUTERHSPBG with the reference's complete argument list (QuatFort.h:90-110), including the eta energy
    well of the three-phase models (2d/quatrhs.m4:356-372)"""
    import ctypes as C
    from ampe_b200 import lib
    from oracle import pyoracle
    L, O = lib.load(), pyoracle.lib()
    n = (13, 9) if ndim == 2 else (10, 7, 6)
    rng = np.random.default_rng(5 + ndim + three_phase)
    cell = tuple(reversed([v + 2 for v in n]))        # ghost width 1 arrays
    cell0 = tuple(reversed(n))
    phi, eta, temp = rng.random(cell), rng.random(cell), 900.0 + 50.0 * rng.random(cell)
    ogm = rng.random(cell0)
    shp = _side_shapes(n)
    flux = [rng.standard_normal(s) for s in shp]
    dx = (C.c_double * 3)(0.2, 0.25, 0.3)
    lo = (C.c_int * 3)(0, 0, 0)
    hi = (C.c_int * 3)(*([v - 1 for v in n] + [0] * (3 - ndim)))
    ref = np.zeros(cell0)
    pf = (C.c_void_p * 3)(*[a.ctypes.data for a in flux])
    ci, vp, dbl, ch = C.c_int, C.c_void_p, C.c_double, C.c_char
    O.oracle_k_computerhspbg.argtypes = [ci, vp, vp, vp, dbl, dbl, vp, ci, vp, ci, dbl, dbl, vp, ci, vp, ci, vp, ci,
                                         vp, ci, ch, ch, ch, ch, ch, ci, ci]
    O.oracle_k_computerhspbg.restype = None
    O.oracle_k_computerhspbg(ndim, lo, hi, dx, 1.7, 0.3, pf, 0, temp.ctypes.data, 1, 2.5, 1.3, phi.ctypes.data, 1,
                             eta.ctypes.data, 1, ogm.ctypes.data, 0, ref.ctypes.data, 0, b"d", b"s", b"p", b"q", b"p",
                             1, three_phase)
    d = lambda a: torch.as_tensor(a).cuda()
    dphi, deta, dtemp, dogm = d(phi), d(eta), d(temp), d(ogm)
    dflux = [d(a) for a in flux]
    got = torch.zeros(cell0, dtype=torch.float64).cuda()
    pg = (C.c_void_p * 3)(*[f.data_ptr() for f in dflux])
    cp = C.c_char_p
    L.ampe_k_computerhspbg.argtypes = [ci, vp, vp, vp, dbl, dbl, vp, ci, vp, ci, dbl, dbl, vp, ci, vp, ci, vp, ci, vp,
                                       ci, cp, cp, cp, cp, cp, ci, ci, vp]
    rc = L.ampe_k_computerhspbg(ndim, lo, hi, dx, 1.7, 0.3, pg, 0, dtemp.data_ptr(), 1, 2.5, 1.3, dphi.data_ptr(), 1,
                                deta.data_ptr() if three_phase else None, 1, dogm.data_ptr(), 0, got.data_ptr(), 0,
                                b"d", b"s", b"p", b"q", b"p", 1, three_phase, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert np.abs(got.cpu().numpy() - ref).max() <= 1e-13 * np.abs(ref).max()
    # three_phase without eta is an argument error, as is an unknown well type
    assert L.ampe_k_computerhspbg(ndim, lo, hi, dx, 1.7, 0.3, pg, 0, dtemp.data_ptr(), 1, 2.5, 1.3, dphi.data_ptr(),
                                  1, None, 1, dogm.data_ptr(), 0, got.data_ptr(), 0, b"d", b"s", b"p", b"q", b"p", 1,
                                  1, None) != 0
