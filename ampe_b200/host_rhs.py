"""ctypes access to the C++ host-side Strategy mirror (ampe_b200/host/*.h,
libampe_b200_host.so): QuatIntegrator::evaluateRHSFunction either through the fused
path or through the reference's own sequence of Strategy calls (piecewise kernels)."""
import ctypes as C
import os

import torch

from . import _abi
from .lib import AmpeError, load

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libampe_b200_host.so")
_lib = None


SUM_REDUCTION = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)  # double (*)(double local, void* user)


def load_host():
    global _lib
    if _lib is None:
        load()  # libampe_b200.so first (dependency, resolved through rpath $ORIGIN)
        if not os.path.exists(_PATH):
            raise AmpeError("libampe_b200_host.so is not built (run __graft_entry__.build())")
        L = C.CDLL(_PATH)
        vp = C.c_void_p
        L.ampe_host_create.restype = vp
        L.ampe_host_create.argtypes = [C.POINTER(_abi.RhsConfig), C.c_int]
        L.ampe_host_destroy.argtypes = [vp]
        L.ampe_host_last_error.restype = C.c_char_p
        L.ampe_host_reset_ref_phase_concentrations.restype = C.c_int
        L.ampe_host_reset_ref_phase_concentrations.argtypes = [vp, vp, vp]
        L.ampe_host_set_symmetry_rotations.restype = C.c_int
        L.ampe_host_set_symmetry_rotations.argtypes = [vp, C.POINTER(vp)]
        L.ampe_host_evaluate_rhs_function.restype = C.c_int
        L.ampe_host_evaluate_rhs_function.argtypes = [vp, C.c_double, C.POINTER(_abi.RhsFields),
                                                      C.POINTER(_abi.RhsFields), C.c_int]
        L.ampe_host_halo_export.restype = C.c_int
        L.ampe_host_halo_export.argtypes = [vp, vp]
        L.ampe_host_halo_connect.restype = C.c_int
        L.ampe_host_halo_connect.argtypes = [vp, vp, vp]
        L.ampe_host_set_sum_reduction.restype = None
        L.ampe_host_set_sum_reduction.argtypes = [vp, SUM_REDUCTION, vp]
        L.ampe_host_integrate_implicit.restype = C.c_int
        L.ampe_host_integrate_implicit.argtypes = [vp, C.POINTER(_abi.RhsFields), C.c_double, C.c_double, C.c_int,
                                                   vp, vp, vp]
        L.ampe_host_integrate_adaptive.restype = C.c_int
        L.ampe_host_integrate_adaptive.argtypes = [vp, C.POINTER(_abi.RhsFields), C.c_double, C.c_double, C.c_double,
                                                   vp, vp, vp]
        L.ampe_host_set_preconditioner.restype = C.c_int
        L.ampe_host_set_preconditioner.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.ampe_host_precond_dquatdphi.restype = C.c_int
        L.ampe_host_precond_dquatdphi.argtypes = [vp, vp, vp]
        L.ampe_host_precond_set.restype = C.c_int
        L.ampe_host_precond_set.argtypes = [vp, C.c_double, C.POINTER(_abi.RhsFields), C.c_double]
        L.ampe_host_precond_solve.restype = C.c_int
        L.ampe_host_precond_solve.argtypes = [vp, C.POINTER(_abi.RhsFields), C.POINTER(_abi.RhsFields)]
        L.ampe_host_precond_level_solver.restype = vp
        L.ampe_host_precond_level_solver.argtypes = [vp, C.c_int]
        L.ampe_host_precond_stats.argtypes = [vp, vp]
        L.ampe_host_read_initial_conditions.restype = C.c_int
        L.ampe_host_read_initial_conditions.argtypes = [C.c_char_p, C.POINTER(_abi.RhsConfig), C.c_int, C.c_int,
                                                        C.POINTER(_abi.RhsFields)]
        L.ampe_host_hdf5_var_shape.restype = C.c_int
        L.ampe_host_hdf5_var_shape.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]
        L.ampe_host_hdf5_read_var.restype = C.c_int
        L.ampe_host_hdf5_read_var.argtypes = [C.c_char_p, C.c_char_p, vp]
        _lib = L
    return _lib


def read_hdf5_variable(filename, name):
    """a whole float / double dataset (rank <= 3) of the root group of an HDF5 / NetCDF-4 file as float64, through the
    HDF5 subset reader of the initial-condition path (host/NetCDF4File.h)"""
    import numpy as np
    L = load_host()
    rank = C.c_int()
    shape = (C.c_longlong * 8)()
    if L.ampe_host_hdf5_var_shape(os.fsencode(filename), name.encode(), C.byref(rank), shape) != 0:
        raise AmpeError(L.ampe_host_last_error().decode())
    out = np.zeros(tuple(shape[:rank.value]), dtype=np.float64)
    if L.ampe_host_hdf5_read_var(os.fsencode(filename), name.encode(), out.ctypes.data) != 0:
        raise AmpeError(L.ampe_host_last_error().decode())
    return out


def read_initial_conditions(filename, cfg, slice_index=-1, fields=("phase", "temperature", "quat", "conc"),
                            device=None):
    """FieldsInitializer::initializeLevelFromData (source/FieldsInitializer.cc:80-360): the state vector of this
    rank's slab from a NetCDF file, classic or NetCDF-4 / HDF5 container (`phase`, `quat1..`, `concentration[0]`, `temperature`, dims z, y, x).
    Returns a dict of float64 tensors in ghost-0 SAMRAI order (pinned host memory when CUDA is available, moved
    to `device` if given); components the configuration does not have are None."""
    L = load_host()
    nz = cfg.n[2] if cfg.ndim == 3 else 1
    shape = (nz, cfg.n[1], cfg.n[0])
    pin = torch.cuda.is_available()
    out = {"phase": None, "quat": None, "conc": None, "temperature": None}
    if cfg.with_phase and "phase" in fields:
        out["phase"] = torch.zeros(shape, dtype=torch.float64, pin_memory=pin)
    if cfg.qlen > 0 and "quat" in fields:
        out["quat"] = torch.zeros((cfg.qlen,) + shape, dtype=torch.float64, pin_memory=pin)
    if cfg.with_concentration and "conc" in fields:
        out["conc"] = torch.zeros(shape, dtype=torch.float64, pin_memory=pin)
    if cfg.with_unsteady_temperature and "temperature" in fields:
        out["temperature"] = torch.zeros(shape, dtype=torch.float64, pin_memory=pin)
    f = _abi.RhsFields()
    for k in out:
        setattr(f, "phase" if k == "phase" else k, None if out[k] is None else out[k].data_ptr())
    mask = (1 if "phase" in fields else 0) | (2 if "temperature" in fields else 0) | (4 if "quat" in fields else 0) | \
        (8 if "conc" in fields else 0)
    rc = L.ampe_host_read_initial_conditions(os.fsencode(filename), C.byref(cfg), int(slice_index), mask, C.byref(f))
    if rc != 0:
        raise AmpeError(L.ampe_host_last_error().decode())
    if device is not None:
        out = {k: (None if v is None else v.to(device, non_blocking=True)) for k, v in out.items()}
    return out


class HostQuatIntegrator:
    def __init__(self, cfg, use_fused):
        self.L = load_host()
        self.h = self.L.ampe_host_create(C.byref(cfg), 1 if use_fused else 0)
        if not self.h:
            raise AmpeError(self.L.ampe_host_last_error().decode())

    def _chk(self, rc):
        if rc != 0:
            raise AmpeError(self.L.ampe_host_last_error().decode())

    def connectSlabRanks(self, rank, nranks, group=None):
        """cfg.nranks > 1: map the neighbours' receive buffers (ampe_halo_* behind QuatIntegrator::haloExport /
        haloConnect; the 128-byte handles travel through torch.distributed here, MPI_Sendrecv in AMPE) and install
        the sum reduction of the vector operations (SAMRAI_MPI::sumReduction in AMPE).  Collective."""
        import torch.distributed as dist
        nb = 128
        mine = C.create_string_buffer(nb)
        self._chk(self.L.ampe_host_halo_export(self.h, mine))
        blobs = [None] * nranks
        dist.all_gather_object(blobs, bytes(mine.raw), group=group)
        prev = C.create_string_buffer(blobs[(rank - 1) % nranks], nb)
        nxt = C.create_string_buffer(blobs[(rank + 1) % nranks], nb)
        rc = self.L.ampe_host_halo_connect(self.h, prev, nxt)
        ok = [None] * nranks
        dist.all_gather_object(ok, int(rc), group=group)
        if any(v != 0 for v in ok):
            self._chk(rc)
            raise AmpeError("a neighbour could not map the receive buffers (codes %r)" % (ok,))
        on_gpu = dist.get_backend(group) == "nccl"

        def _sum(local, _user):
            t = torch.tensor([local], dtype=torch.float64, device="cuda" if on_gpu else "cpu")
            dist.all_reduce(t, group=group)
            return float(t.item())
        self._sum_cb = SUM_REDUCTION(_sum)  # keep the trampoline alive as long as the integrator
        self.L.ampe_host_set_sum_reduction(self.h, self._sum_cb, None)
        dist.barrier(group=group)

    def resetRefPhaseConcentrations(self, cl=None, ca=None):
        self._chk(self.L.ampe_host_reset_ref_phase_concentrations(
            self.h, None if cl is None else cl.data_ptr(), None if ca is None else ca.data_ptr()))

    def setSymmetryRotations(self, iqrot):
        self._iq = [t.to(torch.int32).contiguous() for t in iqrot]
        arr = (C.c_void_p * 3)()
        for d, t in enumerate(self._iq):
            arr[d] = t.data_ptr()
        self._chk(self.L.ampe_host_set_symmetry_rotations(self.h, arr))

    def evaluateRHSFunction(self, time, y, y_dot, fd_flag=0):
        fy, fd = y.fields(), y_dot.fields()
        self._chk(self.L.ampe_host_evaluate_rhs_function(self.h, float(time), C.byref(fy), C.byref(fd),
                                                         int(fd_flag)))
        torch.cuda.synchronize()
        return 0

    IMPLICIT_ENEWTON = -20

    def integrateImplicit(self, y, dt, nsteps, t0=0.0, order=2, max_krylov=5, max_newton=3, rtol=3e-6,
                          atol=3e-4, newton_tol=0.1, lin_factor=0.05):
        """nsteps fixed BDF steps on the device (host/ImplicitIntegrator.h: Newton + matrix-free GMRES with
        fd_flag = 1 Jacobian-vector products, applyProjection after every step), y updated in place.
        Defaults are AMPE's integrator defaults (QuatIntegrator.cc:285-301).  Returns (rc, stats):
        rc 0, or IMPLICIT_ENEWTON when the Newton iteration of a step did not converge."""
        iopt = (C.c_int * 3)(int(order), int(max_krylov), int(max_newton))
        dopt = (C.c_double * 4)(float(rtol), float(atol), float(newton_tol), float(lin_factor))
        st = (C.c_double * 8)()
        fy = y.fields()
        rc = self.L.ampe_host_integrate_implicit(self.h, C.byref(fy), float(t0), float(dt), int(nsteps), iopt,
                                                 dopt, st)
        if rc not in (0, self.IMPLICIT_ENEWTON):
            raise AmpeError(self.L.ampe_host_last_error().decode())
        names = ("steps", "rhs_evals", "jtimes_evals", "newton_iterations", "linear_iterations", "projections",
                 "last_newton_update", "last_linear_residual")
        return rc, dict(zip(names, list(st)))

    ADAPTIVE_STATS = ("steps", "rhs_evals", "jtimes_evals", "newton_iterations", "linear_iterations", "projections",
                      "last_newton_update", "last_linear_residual", "error_test_failures", "convergence_failures",
                      "last_step", "smallest_step", "largest_step", "last_error_estimate", "t_reached")

    def integrateAdaptive(self, y, tend, h0, t0=0.0, order=2, max_krylov=5, max_newton=3, rtol=3e-6, atol=3e-4,
                          newton_tol=0.1, lin_factor=0.05, h_min=0.0, h_max=0.0, max_steps=500, stop_at_tend=True,
                          strict_linear=False, scale_newton_tolerance=False, hold_step_after_failure=False):
        """variable-step BDF1/BDF2 with CVODE's local error test and step controller from t0 to tend on the device
        (host/ImplicitIntegrator.h advanceTo), y updated in place.  Returns (rc, stats); rc 0 or IMPLICIT_E*
        (-20 Newton, -22 too much work, -23 error test, -24 convergence).  stop_at_tend=False: the step sizes are the
        controller's own and the call returns after the first step at or beyond tend (stats["t_reached"] >= tend), the
        way AMPE's run loop meets its output times (one CVODE step per Advance)."""
        iopt = (C.c_int * 5)(int(order), int(max_krylov), int(max_newton), int(max_steps),
                             (0 if stop_at_tend else 1) | (2 if strict_linear else 0) | (4 if scale_newton_tolerance else 0) | (8 if hold_step_after_failure else 0))
        dopt = (C.c_double * 6)(float(rtol), float(atol), float(newton_tol), float(lin_factor), float(h_min),
                                float(h_max))
        st = (C.c_double * 16)()
        fy = y.fields()
        rc = self.L.ampe_host_integrate_adaptive(self.h, C.byref(fy), float(t0), float(tend), float(h0), iopt, dopt,
                                                 st)
        if rc == -1:
            raise AmpeError(self.L.ampe_host_last_error().decode())
        return rc, dict(zip(self.ADAPTIVE_STATS, list(st)))

    # ---- block preconditioners ----
    def setupPreconditioners(self, ncycles=2, precond_has_dquatdphi=False, precondition_left=False):
        """QuatIntegrator::setupPreconditioners: ncycles V-cycles per block solve; 0 = off.  integrateImplicit
        then runs right-preconditioned GMRES (precondition_left: PREC_LEFT like the reference instead).
        precond_has_dquatdphi: with the dquat/dphi coupling block."""
        self._chk(self.L.ampe_host_set_preconditioner(self.h, int(ncycles), 1 if precond_has_dquatdphi else 0,
                                                      1 if precondition_left else 0))

    def multiplyDQuatDPhiBlock(self, phase, qlen):
        """QuatSysSolver::multiplyDQuatDPhiBlock on a ghost-0 CUDA tensor; returns (qlen, ...) tensor"""
        out = torch.empty((qlen,) + tuple(phase.shape[-3:]), dtype=torch.float64, device=phase.device)
        self._chk(self.L.ampe_host_precond_dquatdphi(self.h, phase.data_ptr(), out.data_ptr()))
        return out

    def CVSpgmrPrecondSet(self, t, y, gamma):
        fy = y.fields()
        self._chk(self.L.ampe_host_precond_set(self.h, float(t), C.byref(fy), float(gamma)))

    def CVSpgmrPrecondSolve(self, r, z):
        fr, fz = r.fields(), z.fields()
        self._chk(self.L.ampe_host_precond_solve(self.h, C.byref(fr), C.byref(fz)))

    def preconditionerLevelSolver(self, block):
        """the device multigrid of a block (0 phase, 1 quaternion, 2 composition, 3 temperature) or None"""
        from .precond import LevelSolver
        h = self.L.ampe_host_precond_level_solver(self.h, int(block))
        return LevelSolver(handle=h, owner=self) if h else None

    def precondStats(self):
        out = (C.c_double * 2)()
        self.L.ampe_host_precond_stats(self.h, out)
        return {"precond_setups": out[0], "precond_solves": out[1]}

    def close(self):
        if self.h:
            self.L.ampe_host_destroy(self.h)
            self.h = None
