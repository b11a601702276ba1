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
    for use_fused in (False, True):
        g = _run(cfg, st, use_fused, fds, rot)
        for n, ((status, yo), yg) in enumerate(zip(o_outs, g)):
            for k in ("phase", "quat", "conc", "temperature"):
                if yo.get(k) is None or (k == "quat" and not cfg.evolve_quat):
                    continue
                tol = 1e-11 if (k == "conc" and cfg.free_energy == 2) else parity.TOL
                err = parity.rel_err(yg[k], yo[k])
                assert err <= tol, (name, use_fused, fds[n], k, err)


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
