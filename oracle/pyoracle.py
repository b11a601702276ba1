"""TEST INFRASTRUCTURE ONLY -- ctypes access to the CPU oracle (liboracle.so).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never imported by the product package ampe_b200.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from ampe_b200 import _abi  # noqa: E402  (struct layouts only)


_TARGETS = {False: "liboracle.so", True: "liboracle_perf.so", "par": "liboracle_par.so"}


def build(perf=False):
    target = _TARGETS[perf]
    subprocess.check_call(["make", "-s", "-C", _HERE, target])
    return os.path.join(_HERE, target)


_libs = {}


def _stale(path):
    """a source newer than the library (an edit without a rebuild): let make decide"""
    try:
        t = os.path.getmtime(path)
        srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cc", ".h"))]
        srcs.append(os.path.join(os.path.dirname(_HERE), "ampe_b200", "csrc", "mg_cell.h"))
        return any(os.path.getmtime(f) > t for f in srcs if os.path.exists(f))
    except OSError:
        return False


def lib(perf=False):
    if perf not in _libs:
        path = os.path.join(_HERE, _TARGETS[perf])
        if not os.path.exists(path) or _stale(path):
            build(perf)
        L = C.CDLL(path)
        dbl = C.c_double
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(_abi.RhsConfig)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_set_rotations.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.oracle_eval.restype = C.c_int
        L.oracle_eval.argtypes = [C.c_void_p, dbl, C.POINTER(_abi.RhsFields),
                                  C.POINTER(_abi.RhsFields), C.c_int]
        L.oracle_get_phase_concentrations.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_energy.restype = C.c_int
        L.oracle_energy.argtypes = [C.c_void_p, C.POINTER(_abi.RhsFields), C.c_void_p]
        L.oracle_scalar_diagnostics.restype = C.c_int
        L.oracle_scalar_diagnostics.argtypes = [C.c_void_p, C.POINTER(_abi.RhsFields), C.c_void_p]
        for f in ("interp_func", "deriv_interp_func", "second_deriv_interp_func", "well_func",
                  "deriv_well_func"):
            fn = getattr(L, "oracle_" + f)
            fn.restype = dbl
            fn.argtypes = [dbl, C.c_char]
        L.oracle_average_func.restype = dbl
        L.oracle_average_func.argtypes = [dbl, dbl, C.c_char]
        for f in ("interp_ratio_func", "compl_interp_ratio_func"):
            fn = getattr(L, "oracle_" + f)
            fn.restype = dbl
            fn.argtypes = [dbl, C.c_char, C.c_char]
        L.oracle_eval_grad_normi.restype = dbl
        L.oracle_eval_grad_normi.argtypes = [dbl, C.c_char, dbl, dbl]
        L.oracle_quatsymmrotate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_qr_table4.argtypes = [C.c_void_p]
        L.oracle_integrate_implicit.restype = C.c_int
        L.oracle_integrate_implicit.argtypes = [C.c_void_p, C.POINTER(_abi.RhsFields), dbl, dbl, C.c_int,
                                                C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_integrate_adaptive.restype = C.c_int
        L.oracle_integrate_adaptive.argtypes = [C.c_void_p, C.POINTER(_abi.RhsFields), dbl, dbl, dbl,
                                                C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_quatfindsymm.restype = C.c_int
        L.oracle_quatfindsymm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_k_quat_symm_rotation.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                  C.c_int, C.POINTER(C.c_void_p), C.c_int]
        L.oracle_k_quat_fundamental.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_k_project.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        pdb = C.POINTER(_abi.CalphadBinary)
        for f in ("calphad_free_energy", "calphad_deriv_free_energy",
                  "calphad_second_deriv_free_energy"):
            fn = getattr(L, "oracle_" + f)
            fn.restype = dbl
            fn.argtypes = [pdb, dbl, dbl, C.c_int]
        L.oracle_calphad_phase_concentrations.restype = C.c_int
        L.oracle_calphad_phase_concentrations.argtypes = [pdb, dbl, dbl, dbl, C.c_void_p, dbl,
                                                          C.c_int, dbl]
        L.oracle_calphad_ceq.restype = C.c_int
        L.oracle_calphad_ceq.argtypes = [pdb, dbl, C.c_void_p, dbl, C.c_int, dbl]
        for f in ("calphad_fmix", "calphad_fmix_deriv", "calphad_fmix_deriv2"):
            fn = getattr(L, "oracle_" + f)
            fn.restype = dbl
            fn.argtypes = [dbl] * 5
        for f in ("xlogx", "xlogx_deriv", "xlogx_deriv2"):
            fn = getattr(L, "oracle_" + f)
            fn.restype = dbl
            fn.argtypes = [dbl]
        L.oracle_calphad_diffusion_mobility.restype = dbl
        L.oracle_calphad_diffusion_mobility.argtypes = [pdb, C.c_int, dbl, dbl]
        # block preconditioners (precond.cc)
        vp, pvp, ci = C.c_void_p, C.POINTER(C.c_void_p), C.c_int
        L.oracle_set_preconditioner.argtypes = [vp, ci, ci, ci]
        L.oracle_precond_dquatdphi.restype = ci
        L.oracle_precond_dquatdphi.argtypes = [vp, vp, vp]
        L.oracle_precond_stats.argtypes = [vp, vp]
        L.oracle_precond_setup.restype = ci
        L.oracle_precond_setup.argtypes = [vp, dbl, ci, ci]
        L.oracle_precond_solve.restype = ci
        L.oracle_precond_solve.argtypes = [vp, C.POINTER(_abi.RhsFields), C.POINTER(_abi.RhsFields)]
        L.oracle_precond_apply.restype = ci
        L.oracle_precond_apply.argtypes = [vp, ci, vp, vp]
        L.oracle_precond_block.restype = vp
        L.oracle_precond_block.argtypes = [vp, ci]
        L.oracle_mg_create.restype = vp
        L.oracle_mg_create.argtypes = [ci, vp, vp, ci, ci]
        L.oracle_mg_num_components.restype = ci
        L.oracle_mg_num_components.argtypes = [vp]
        L.oracle_mg_destroy.argtypes = [vp]
        L.oracle_mg_set_elliptic.restype = ci
        L.oracle_mg_set_elliptic.argtypes = [vp, vp, ci, dbl, vp, ci, dbl, pvp, pvp, ci, dbl, dbl]
        L.oracle_mg_set_quat.restype = ci
        L.oracle_mg_set_quat.argtypes = [vp, dbl, vp, ci, pvp, ci]
        L.oracle_mg_solve.restype = ci
        L.oracle_mg_solve.argtypes = [vp, vp, vp, ci, ci]
        L.oracle_mg_apply.argtypes = [vp, vp, vp]
        L.oracle_mg_set_sweeps.argtypes = [vp, ci, ci, ci]
        L.oracle_mg_set_zero_slope.argtypes = [vp, vp]
        L.oracle_mg_set_fused.restype = ci
        L.oracle_mg_set_fused.argtypes = [vp, ci, C.c_longlong]
        L.oracle_mg_num_levels.restype = ci
        L.oracle_mg_num_levels.argtypes = [vp]
        L.oracle_mg_level_extents.argtypes = [vp, ci, vp]
        L.oracle_mg_copy_level.restype = ci
        L.oracle_mg_copy_level.argtypes = [vp, ci, ci, vp]
        L.oracle_k_elliptic_apply.argtypes = [ci, vp, vp, vp, ci, vp, ci, pvp, vp, vp]
        L.oracle_k_quat_stencil_apply.argtypes = [ci, vp, vp, dbl, vp, ci, pvp, vp, vp]
        _libs[perf] = L
    return _libs[perf]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _fields(d):
    f = _abi.RhsFields()
    f.phase = _ptr(d.get("phase"))
    f.quat = _ptr(d.get("quat"))
    f.conc = _ptr(d.get("conc"))
    f.temperature = _ptr(d.get("temperature"))
    return f


class Oracle:
    """CPU restatement of QuatIntegrator::evaluateRHSFunction on one periodic level."""

    def __init__(self, cfg, perf=False):
        self.L = lib(perf)
        self.cfg = cfg
        self.h = self.L.oracle_create(C.byref(cfg))
        self.ncell = cfg.n[0] * cfg.n[1] * (cfg.n[2] if cfg.ndim == 3 else 1)

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_ref(self, cl, ca):
        self.L.oracle_set_ref(self.h, _ptr(cl), _ptr(ca))

    def set_rotations(self, iqrot):
        arr = (C.c_void_p * 3)()
        self._iq = [np.ascontiguousarray(a, dtype=np.int32) for a in iqrot]
        for d, a in enumerate(self._iq):
            arr[d] = a.ctypes.data
        self.L.oracle_set_rotations(self.h, arr)

    def alloc_like(self, y):
        return {k: (None if v is None else np.zeros_like(v)) for k, v in y.items()}

    def eval(self, time, y, fd_flag=0, ydot=None):
        """y: dict of contiguous float64 numpy arrays (ghost-0 SAMRAI order)."""
        if ydot is None:
            ydot = self.alloc_like(y)
        fy, fd = _fields(y), _fields(ydot)
        st = self.L.oracle_eval(self.h, float(time), C.byref(fy), C.byref(fd), int(fd_flag))
        return st, ydot

    def energy(self, y):
        """QuatModel::evaluateEnergy: (status, [total, phi, orient, qint, well, free, 0, 0])"""
        out = np.zeros(8)
        st = self.L.oracle_energy(self.h, C.byref(_fields(y)), _ptr(out))
        return st, out

    DIAGNOSTICS = ("volume", "volume_solid", "solid_fraction", "integral_concentration", "max_concentration",
                   "integral_phase_concentration", "cex", "min_temperature", "max_temperature",
                   "average_temperature", "thermal_energy")

    def grain_volumes(self, y, phase_threshold=0.85, max_grains=4096, numbers=False):
        """Grains::findAndNumberGrains + computeGrainVolumes: {grain number: volume} (and the per-cell numbers)"""
        cfg = self.cfg
        n = np.array([cfg.n[0], cfg.n[1], cfg.n[2]], dtype=np.int32)
        dx = np.array([cfg.dx[0], cfg.dx[1], cfg.dx[2]], dtype=np.float64)
        zs = np.array([cfg.zero_slope[0], cfg.zero_slope[1], cfg.zero_slope[2]], dtype=np.int32)
        phase = np.ascontiguousarray(y["phase"], dtype=np.float64)
        ids = np.zeros(max_grains, dtype=np.int32)
        vols = np.zeros(max_grains)
        ng = C.c_int(0)
        num = np.zeros(phase.size, dtype=np.int32) if numbers else None
        fn = self.L.oracle_grain_volumes
        fn.restype = C.c_int
        fn.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_double, C.c_int] + [C.c_void_p] * 4
        rc = fn(int(cfg.ndim), _ptr(n), _ptr(dx), _ptr(zs), _ptr(phase), float(phase_threshold), int(max_grains),
                C.addressof(ng), _ptr(ids), _ptr(vols), None if num is None else _ptr(num))
        assert rc == 0, "more than max_grains grains (%d)" % ng.value
        out = {int(ids[i]): float(vols[i]) for i in range(ng.value)}
        return (out, num.reshape(phase.shape)) if numbers else out

    def scalar_diagnostics(self, y):
        """QuatModel::printScalarDiagnostics: dict of the DIAGNOSTICS"""
        out = np.zeros(12)
        assert self.L.oracle_scalar_diagnostics(self.h, C.byref(_fields(y)), _ptr(out)) == 0
        return dict(zip(self.DIAGNOSTICS, out.tolist()))

    def integrate_implicit(self, y, dt, nsteps, t0=0.0, order=2, max_krylov=5, max_newton=3, rtol=3e-6,
                           atol=3e-4, newton_tol=0.1, lin_factor=0.05):
        """the product's ImplicitIntegrator template (ampe_b200/host/ImplicitIntegrator.h) driven by the
        oracle's RHS: y (dict of numpy arrays) is advanced in place; returns (rc, stats dict)"""
        iopt = np.array([order, max_krylov, max_newton], dtype=np.int32)
        dopt = np.array([rtol, atol, newton_tol, lin_factor], dtype=np.float64)
        st = np.zeros(8)
        fy = _fields(y)
        rc = self.L.oracle_integrate_implicit(self.h, C.byref(fy), float(t0), float(dt), int(nsteps),
                                              _ptr(iopt), _ptr(dopt), _ptr(st))
        names = ("steps", "rhs_evals", "jtimes_evals", "newton_iterations", "linear_iterations", "projections",
                 "last_newton_update", "last_linear_residual")
        return rc, dict(zip(names, st.tolist()))

    # ---- block preconditioners (precond.cc) ----
    def set_preconditioner(self, ncycles, dquatdphi=False, left=False):
        """ncycles > 0: integrate_implicit runs right-preconditioned GMRES (CVSpgmrPrecondSet / Solve);
        dquatdphi: with the lower-triangular dquat/dphi coupling block (precond_has_dquatdphi); left: PREC_LEFT like
        the reference instead of right preconditioning"""
        self.L.oracle_set_preconditioner(self.h, int(ncycles), 1 if dquatdphi else 0, 1 if left else 0)

    def precond_stats(self):
        out = np.zeros(2)
        self.L.oracle_precond_stats(self.h, _ptr(out))
        return {"precond_setups": out[0], "precond_solves": out[1]}

    def precond_setup(self, gamma, ncycles=2, dquatdphi=False):
        """coefficients frozen at the state of the last fd_flag = 0 eval (CVSpgmrPrecondSet)"""
        return self.L.oracle_precond_setup(self.h, float(gamma), int(ncycles), 1 if dquatdphi else 0)

    def precond_dquatdphi(self, z_phase):
        """QuatFACOps::multiplyDQuatDPhiBlock: [dF_q/dphi] z_phase, shape (qlen, ...)"""
        z = np.ascontiguousarray(z_phase, dtype=np.float64)
        out = np.zeros((self.cfg.qlen,) + z.shape[-3:] if z.ndim >= 3 else (self.cfg.qlen, z.size))
        if self.L.oracle_precond_dquatdphi(self.h, _ptr(z), _ptr(out)) != 0:
            raise RuntimeError("oracle_precond_dquatdphi: block not set up")
        return out

    def precond_solve(self, r, z=None):
        if z is None:
            z = self.alloc_like(r)
        fr, fz = _fields(r), _fields(z)
        return self.L.oracle_precond_solve(self.h, C.byref(fr), C.byref(fz)), z

    def precond_apply(self, block, u):
        """restated reference operator of a block (0 phase, 1 quaternion component, 2 composition,
        3 temperature) applied to one ghost-0 cell array"""
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.zeros_like(u)
        rc = self.L.oracle_precond_apply(self.h, int(block), _ptr(u), _ptr(out))
        if rc != 0:
            raise RuntimeError("oracle_precond_apply: no such block")
        return out

    def precond_block(self, block):
        h = self.L.oracle_precond_block(self.h, int(block))
        return HostMG(handle=h, owner=self) if h else None

    ADAPTIVE_STATS = ("steps", "rhs_evals", "jtimes_evals", "newton_iterations", "linear_iterations", "projections",
                      "last_newton_update", "last_linear_residual", "error_test_failures", "convergence_failures",
                      "last_step", "smallest_step", "largest_step", "last_error_estimate", "t_reached")

    def integrate_adaptive(self, y, tend, h0, t0=0.0, order=2, max_krylov=5, max_newton=3, rtol=3e-6, atol=3e-4,
                           newton_tol=0.1, lin_factor=0.05, h_min=0.0, h_max=0.0, max_steps=500, stop_at_tend=True,
                           strict_linear=False, scale_newton_tolerance=False, hold_step_after_failure=False):
        """ImplicitIntegrator::advanceTo driven by the oracle's RHS: variable steps with the local error test
        from t0 to tend; y advanced in place; returns (rc, stats)"""
        iopt = np.array([order, max_krylov, max_newton, max_steps, (0 if stop_at_tend else 1) | (2 if strict_linear else 0) | (4 if scale_newton_tolerance else 0) | (8 if hold_step_after_failure else 0)],
                        dtype=np.int32)
        dopt = np.array([rtol, atol, newton_tol, lin_factor, h_min, h_max], dtype=np.float64)
        st = np.zeros(16)
        fy = _fields(y)
        rc = self.L.oracle_integrate_adaptive(self.h, C.byref(fy), float(t0), float(tend), float(h0), _ptr(iopt),
                                              _ptr(dopt), _ptr(st))
        return rc, dict(zip(self.ADAPTIVE_STATS, st.tolist()))

    def phase_concentrations(self):
        cl = np.zeros(self.ncell)
        ca = np.zeros(self.ncell)
        self.L.oracle_get_phase_concentrations(self.h, _ptr(cl), _ptr(ca))
        return cl, ca


def _ivec(v):
    return (C.c_int * 3)(*(list(v) + [0] * (3 - len(v))))


def _parr(arrays):
    """host array of pointers (NULL-padded to 3) from a list of numpy arrays, or None"""
    if arrays is None:
        return None
    arr = (C.c_void_p * 3)()
    for d, a in enumerate(arrays):
        assert a.dtype == np.float64 and a.flags.c_contiguous
        arr[d] = a.ctypes.data
    return arr


class HostMG:
    """host loop over the product's per-cell multigrid functions (oracle/precond.cc part 2): the same
    calls as ampe_b200's device solver (ampe_mg_*), on numpy arrays"""

    def __init__(self, n=None, dx=None, with_s=False, handle=None, owner=None, ncomp=1):
        """ncomp components solved together with one matrix (rhs / solution arrays of shape (ncomp, nz, ny, nx))"""
        self.L = lib()
        self._owner = owner  # borrowed handle of an Oracle context
        if handle is not None:
            self.h, self._own = handle, False
        else:
            self.h = self.L.oracle_mg_create(len(n), _ivec(n), (C.c_double * 3)(*(list(dx) + [0.0] * (3 - len(dx)))),
                                             1 if with_s else 0, int(ncomp))
            self._own = True
        self._keep = []

    def __del__(self):
        if getattr(self, "_own", False) and self.h:
            self.L.oracle_mg_destroy(self.h)
            self.h = None

    def set_elliptic(self, m=None, ngm=0, m_const=0.0, c=None, ngc=0, c_const=0.0, d=None, d2=None, ngd=0,
                     d_scale=1.0, d_const=0.0):
        rc = self.L.oracle_mg_set_elliptic(self.h, _ptr(m), ngm, m_const, _ptr(c), ngc, c_const, _parr(d), _parr(d2),
                                           ngd, d_scale, d_const)
        assert rc == 0

    def set_quat(self, gamma, mobility, ngm, face_coef, ngfc):
        assert self.L.oracle_mg_set_quat(self.h, gamma, _ptr(mobility), ngm, _parr(face_coef), ngfc) == 0

    def solve(self, rhs, ncycles=2, symmetrized=False):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        out = np.zeros_like(rhs)
        assert self.L.oracle_mg_solve(self.h, _ptr(rhs), _ptr(out), int(ncycles), 1 if symmetrized else 0) == 0
        return out

    def apply(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.zeros_like(u)
        self.L.oracle_mg_apply(self.h, _ptr(u), _ptr(out))
        return out

    def set_sweeps(self, pre, post, coarse):
        self.L.oracle_mg_set_sweeps(self.h, pre, post, coarse)

    def set_zero_slope(self, zero_slope):
        self.L.oracle_mg_set_zero_slope(self.h, _ivec(list(zero_slope)))

    def set_fused(self, on=True, min_cells=4096):
        """one red-black sweep per pass over tiles (mg_rb_tile_pass) on the levels with more than min_cells
        cells; returns the number of levels that use it"""
        return self.L.oracle_mg_set_fused(self.h, 1 if on else 0, int(min_cells))

    def num_levels(self):
        return self.L.oracle_mg_num_levels(self.h)

    def level_extents(self, level):
        n = (C.c_int * 3)()
        self.L.oracle_mg_level_extents(self.h, level, n)
        return list(n)

    def level_array(self, level, which):
        n = self.level_extents(level)
        out = np.zeros((n[2], n[1], n[0]))
        if self.L.oracle_mg_copy_level(self.h, level, which, _ptr(out)) != 0:
            return None
        return out


def elliptic_apply(n, dx, m, ngm, c, ngc, d, u):
    """restated efo_compfluxvardc + efo_compresvarsca: M div(D grad u) + C u (SAMRAI-layout m, c, d)"""
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros_like(u)
    lib().oracle_k_elliptic_apply(len(n), _ivec(n), (C.c_double * 3)(*(list(dx) + [0.0] * (3 - len(dx)))), _ptr(m),
                                  ngm, _ptr(c), ngc, _parr(d), _ptr(u), _ptr(out))
    return out


def quat_stencil_apply(n, dx, gamma, sqrt_m, ngm, fc, w):
    """restated set_j_ij + set_stencil applied to one component"""
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.zeros_like(w)
    lib().oracle_k_quat_stencil_apply(len(n), _ivec(n), (C.c_double * 3)(*(list(dx) + [0.0] * (3 - len(dx)))),
                                      float(gamma), _ptr(sqrt_m), ngm, _parr(fc), _ptr(w), _ptr(out))
    return out


# ---- symmetry pre-pass / projection (oracle/symmetry.cc), SAMRAI layouts ----


def quatfindsymm(q1, q2, iq, qlen):
    """quatfindsymm (quat.f:9-37): (rotation index, rotated q2)"""
    q1 = np.ascontiguousarray(q1, dtype=np.float64)
    q2 = np.ascontiguousarray(q2, dtype=np.float64)
    out = np.zeros(qlen)
    r = lib().oracle_quatfindsymm(_ptr(q1), _ptr(q2), int(iq), _ptr(out), int(qlen))
    return r, out


def quat_symm_rotation(n, q_ghosted, ngq, depth, rot, ngrot):
    """QUAT_SYMM_ROTATION on the box [0, n-1]; q_ghosted: (depth, [nz+2g,] ny+2g, nx+2g);
    rot: list of int32 side arrays (in/out)"""
    ndim = len(n)
    arr = (C.c_void_p * 3)()
    for d, a in enumerate(rot):
        assert a.dtype == np.int32 and a.flags.c_contiguous
        arr[d] = a.ctypes.data
    lib().oracle_k_quat_symm_rotation(ndim, _ivec([0] * ndim), _ivec([v - 1 for v in n]), _ptr(q_ghosted),
                                      int(ngq), int(depth), arr, int(ngrot))


def quat_fundamental(n, q_ghosted, ngq, depth):
    ndim = len(n)
    lib().oracle_k_quat_fundamental(ndim, _ivec([0] * ndim), _ivec([v - 1 for v in n]), _ptr(q_ghosted),
                                    int(ngq), int(depth))


def project(n, depth, q, ngq, corr, ngc, err, nge):
    ndim = len(n)
    lib().oracle_k_project(ndim, _ivec([0] * ndim), _ivec([v - 1 for v in n]), int(depth), _ptr(q), int(ngq),
                           _ptr(corr), int(ngc), _ptr(err), int(nge))


# ---- extended-precision arbiter (liboracle_ld.so: the same sources with double -> long double) ------------------
_LD_CLASSES = {}


def _ld_type(t):
    """ctypes type with every c_double replaced by c_longdouble (structures and arrays recursively)"""
    if t is C.c_double:
        return C.c_longdouble
    if isinstance(t, type) and issubclass(t, C.Array):
        return _ld_type(t._type_) * t._length_
    if isinstance(t, type) and issubclass(t, C.Structure):
        if t not in _LD_CLASSES:
            _LD_CLASSES[t] = type(t.__name__ + "LD", (C.Structure,),
                                  {"_fields_": [(n, _ld_type(ft)) for n, ft in t._fields_]})
        return _LD_CLASSES[t]
    return t


def _ld_copy(src, dst):
    for name, ft in type(src)._fields_:
        v = getattr(src, name)
        if isinstance(v, C.Array):
            _ld_copy_array(v, getattr(dst, name))
        elif isinstance(v, C.Structure):
            _ld_copy(v, getattr(dst, name))
        else:
            setattr(dst, name, v)


def _ld_copy_array(src, dst):
    for i in range(len(src)):
        if isinstance(src[i], C.Array):
            _ld_copy_array(src[i], dst[i])
        elif isinstance(src[i], C.Structure):
            _ld_copy(src[i], dst[i])
        else:
            dst[i] = src[i]


def lib_ld():
    if "ld" not in _libs:
        path = os.path.join(_HERE, "liboracle_ld.so")
        if not os.path.exists(path) or _stale(path):
            subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_ld.so"])
        L = C.CDLL(path)
        L.oracle_ld_create.restype = C.c_void_p
        L.oracle_ld_create.argtypes = [C.c_void_p]
        L.oracle_ld_destroy.argtypes = [C.c_void_p]
        L.oracle_ld_set_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_ld_set_rotations.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_ld_eval.restype = C.c_int
        L.oracle_ld_eval.argtypes = [C.c_void_p, C.c_longdouble, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_ld_get_phase_concentrations.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        assert L.oracle_ld_sizeof_real() == np.dtype(np.longdouble).itemsize
        _libs["ld"] = L
    return _libs["ld"]


class OracleLD:
    """The restatement evaluated in long double (64-bit mantissa): the arbiter for outputs that two correct fp64
    evaluations cannot agree on to 1e-12.  Inputs are the SAME fp64 fields (widened exactly), outputs are returned
    as numpy longdouble arrays."""

    def __init__(self, cfg):
        self.L = lib_ld()
        self.cfg = cfg
        self.cfg_ld = _ld_type(type(cfg))()
        _ld_copy(cfg, self.cfg_ld)
        self.h = self.L.oracle_ld_create(C.byref(self.cfg_ld))
        self.ncell = cfg.n[0] * cfg.n[1] * (cfg.n[2] if cfg.ndim == 3 else 1)

    def close(self):
        if self.h:
            self.L.oracle_ld_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @staticmethod
    def _wide(a):
        return None if a is None else np.ascontiguousarray(a, dtype=np.longdouble)

    def _fields(self, d):
        f = _ld_type(_abi.RhsFields)()
        for k in ("phase", "quat", "conc", "temperature"):
            a = d.get(k)
            setattr(f, k, None if a is None else a.ctypes.data)
        return f

    def set_ref(self, cl, ca):
        self._ref = (self._wide(cl), self._wide(ca))
        self.L.oracle_ld_set_ref(self.h, _ptr(self._ref[0]), _ptr(self._ref[1]))

    def set_rotations(self, iqrot):
        arr = (C.c_void_p * 3)()
        self._iq = [np.ascontiguousarray(a, dtype=np.int32) for a in iqrot]
        for d, a in enumerate(self._iq):
            arr[d] = a.ctypes.data
        self.L.oracle_ld_set_rotations(self.h, arr)

    def eval(self, time, y, fd_flag=0):
        yw = {k: self._wide(v) for k, v in y.items()}
        ydot = {k: (None if v is None else np.zeros_like(v)) for k, v in yw.items()}
        fy, fd = self._fields(yw), self._fields(ydot)
        st = self.L.oracle_ld_eval(self.h, float(time), C.byref(fy), C.byref(fd), int(fd_flag))
        return st, ydot

    def phase_concentrations(self):
        cl = np.zeros(self.ncell, dtype=np.longdouble)
        ca = np.zeros(self.ncell, dtype=np.longdouble)
        self.L.oracle_ld_get_phase_concentrations(self.h, _ptr(cl), _ptr(ca))
        return cl, ca
