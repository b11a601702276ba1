"""Python mirror of the reference's integrator-facing interface for the RHS path.

`QuatIntegratorRHS.evaluateRHSFunction(time, y, y_dot, fd_flag)` has the argument
meaning and error behaviour of QuatIntegrator::evaluateRHSFunction
(source/QuatIntegrator.h:204-222): y is not modified, 0 is returned on success.
y / y_dot are `SolutionVector`s: named ghost-0 components in the order of
createSolutionvector (QuatIntegrator.cc:1623-1674), held as torch CUDA tensors
(torch is used for device memory and streams only; all arithmetic happens in
libampe_b200.so through the C ABI)."""
import ctypes as C

import torch

from . import _abi
from .lib import AmpeError, check, load

COMPONENTS = ("phase", "quat", "conc", "temperature")


class SolutionVector(dict):
    """Stand-in for Sundials_SAMRAIVector on one uniform level: component name ->
    tensor of shape (depth, nz, ny, nx) in SAMRAI CellData order (i fastest)."""

    def fields(self):
        f = _abi.RhsFields()
        for k in COMPONENTS:
            t = self.get(k)
            setattr(f, k, None if t is None else t.data_ptr())
        return f

    def like(self):
        return SolutionVector({k: (None if v is None else torch.zeros_like(v)) for k, v in self.items()})


class QuatIntegratorRHS:
    """One context per GPU/stream; calls on a context are serialised by the caller
    (the reference is not re-entrant either: SURVEY.md 8b 'Threading')."""

    def __init__(self, cfg, device=None):
        if not torch.cuda.is_available():
            raise AmpeError("no CUDA device: ampe_b200 has no CPU fallback")
        self.L = load()
        self.cfg = cfg
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.L.ampe_rhs_create(C.byref(cfg), C.byref(h)), "ampe_rhs_create")
        self.h = h
        self.ncell = cfg.n[0] * cfg.n[1] * (cfg.n[2] if cfg.ndim == 3 else 1)
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.ampe_rhs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # QuatModel::resetRefPhaseConcentrations (QuatModel.cc:5218-5231)
    def resetRefPhaseConcentrations(self, cl_ref=None, ca_ref=None):
        a = None if cl_ref is None else cl_ref.data_ptr()
        b = None if ca_ref is None else ca_ref.data_ptr()
        check(self.L.ampe_rhs_set_ref_concentrations(self.h, a, b, self._stream()), "set_ref")

    def setRefPhaseConcentrationsGhosted(self, cl_g, ca_g):
        check(self.L.ampe_rhs_set_ref_concentrations_ghosted(self.h, cl_g.data_ptr(), ca_g.data_ptr(),
                                                             self._stream()), "set_ref_ghosted")

    def setSymmetryRotations(self, iqrot):
        arr = (C.c_void_p * 3)()
        self._iq = [t.to(torch.int32).contiguous() for t in iqrot]
        for d, t in enumerate(self._iq):
            arr[d] = t.data_ptr()
        check(self.L.ampe_rhs_set_symmetry_rotations(self.h, arr, self._stream()), "set_rotations")

    def setHalo(self, lo, hi):
        """lo / hi: SolutionVector of ghost planes from the lower / upper slab neighbour."""
        if lo is None:
            check(self.L.ampe_rhs_set_halo(self.h, None, None), "set_halo")
            return
        self._halo = (lo, hi)
        flo, fhi = lo.fields(), hi.fields()
        check(self.L.ampe_rhs_set_halo(self.h, C.byref(flo), C.byref(fhi)), "set_halo")

    def nghosts(self):
        return self.L.ampe_rhs_nghosts(self.h)

    def evaluateRHSFunction(self, time, y, y_dot, fd_flag=0, part=0):
        fy, fd = y.fields(), y_dot.fields()
        fn = (self.L.ampe_rhs_eval, self.L.ampe_rhs_eval_interior, self.L.ampe_rhs_eval_boundary)[part]
        check(fn(self.h, float(time), C.byref(fy), C.byref(fd), int(fd_flag), self._stream()),
              "evaluateRHSFunction")
        return 0

    def evaluateRHSFunctionHost(self, time, y_host, ydot_host, fd_flag=0):
        """Reference-facing plugin path: HOST buffers (numpy / pinned torch CPU tensors);
        host->device and device->host copies happen inside the call."""
        fy, fd = _abi.RhsFields(), _abi.RhsFields()
        for k in COMPONENTS:
            a, b = y_host.get(k), ydot_host.get(k)
            setattr(fy, k, None if a is None else a.data_ptr())
            setattr(fd, k, None if b is None else b.data_ptr())
        check(self.L.ampe_rhs_eval_host(self.h, float(time), C.byref(fy), C.byref(fd), int(fd_flag)),
              "evaluateRHSFunctionHost")
        return 0

    # ---- device vector operations, normalizeQuat, fixed-step integrator, energy
    def linearSum(self, a, x, b, y, z):
        """z = a x + b y on the evolved components (Sundials_SAMRAIVector::linearSum)"""
        fx, fy, fz = x.fields(), y.fields(), z.fields()
        check(self.L.ampe_vec_linear_sum(self.h, float(a), C.byref(fx), float(b), C.byref(fy), C.byref(fz),
                                         self._stream()), "linearSum")

    def scale(self, a, x, z):
        fx, fz = x.fields(), z.fields()
        check(self.L.ampe_vec_scale(self.h, float(a), C.byref(fx), C.byref(fz), self._stream()), "scale")

    def _reduce(self, fn, *vecs):
        out = C.c_double(0.0)
        fs = [v.fields() for v in vecs]
        check(fn(self.h, *[C.byref(f) for f in fs], C.byref(out), self._stream()), fn.__name__)
        return out.value

    def dotWith(self, x, y):
        return self._reduce(self.L.ampe_vec_dot, x, y)

    def weightedRMSNorm(self, x, w):
        return self._reduce(self.L.ampe_vec_wrms_norm, x, w)

    def maxNorm(self, x):
        return self._reduce(self.L.ampe_vec_max_norm, x)

    def normalizeQuat(self, y):
        """QuatModel::normalizeQuat (QuatModel.cc:4222-4262)"""
        fy = y.fields()
        check(self.L.ampe_normalize_quat(self.h, C.byref(fy), self._stream()), "normalizeQuat")

    def integrateFixed(self, y, dt, nsteps, scheme=0, t0=0.0):
        """nsteps explicit steps of size dt on the device, y updated in place"""
        w1 = y.like()
        w2 = y.like() if scheme == 1 else None
        fy, f1 = y.fields(), w1.fields()
        f2 = w2.fields() if w2 is not None else None
        check(self.L.ampe_integrate_fixed(self.h, C.byref(fy), C.byref(f1),
                                          C.byref(f2) if f2 is not None else None, float(t0), float(dt),
                                          int(nsteps), int(scheme), self._stream()), "integrateFixed")
        torch.cuda.current_stream().synchronize()  # work vectors are released on return

    def evaluateEnergy(self, y):
        """QuatModel::evaluateEnergy: dict of total / phase / orient / qint / well / free"""
        out = (C.c_double * 8)()
        fy = y.fields()
        check(self.L.ampe_energy_eval(self.h, C.byref(fy), out, self._stream()), "evaluateEnergy")
        names = ("total", "phase", "orient", "qint", "well", "free")
        return {k: out[i] for i, k in enumerate(names)}

    DIAGNOSTICS = ("volume", "volume_solid", "solid_fraction", "integral_concentration", "max_concentration",
                   "integral_phase_concentration", "cex", "min_temperature", "max_temperature",
                   "average_temperature", "thermal_energy")

    def printScalarDiagnostics(self, y):
        """QuatModel::printScalarDiagnostics (QuatModel.cc:2543-2690): the scalars the reference prints every
        d_scalar_diag_interval -- volume fraction of solid, integral / max concentration, Cex, temperature
        extrema and average, thermal energy -- as a dict (this rank's cells)"""
        out = (C.c_double * 12)()
        fy = y.fields()
        check(self.L.ampe_scalar_diagnostics(self.h, C.byref(fy), out, self._stream()), "printScalarDiagnostics")
        return {k: out[i] for i, k in enumerate(self.DIAGNOSTICS)}

    def computeGrainDiagnostics(self, y, phase_threshold=0.85, max_grains=4096, numbers=False):
        """QuatModel::computeGrainDiagnostics (QuatModel.cc:2690-2705): Grains::findAndNumberGrains +
        computeGrainVolumes -- {grain number: volume}, the "Volume of grain N = V" lines of the reference's output;
        numbers=True also returns the per-cell grain numbers (int32 device tensor, -1 outside grains)"""
        ids = (C.c_int * max_grains)()
        vols = (C.c_double * max_grains)()
        n = C.c_int(0)
        fy = y.fields()
        check(self.L.ampe_grain_volumes(self.h, C.byref(fy), float(phase_threshold), int(max_grains), C.byref(n), ids,
                                        vols, self._stream()), "computeGrainDiagnostics")
        out = {int(ids[i]): float(vols[i]) for i in range(n.value)}
        if not numbers:
            return out
        num = torch.empty(y["phase"].shape, dtype=torch.int32, device=y["phase"].device)
        check(self.L.ampe_grain_numbers(self.h, C.c_void_p(num.data_ptr())), "ampe_grain_numbers")
        return out, num

    def applyProjection(self, time, y, corr, epsProj, err):
        """QuatIntegrator::applyProjection (QuatIntegrator.cc:3911-3962): corr <- 0 except the
        quaternion part, where y + corr is normalised; err loses its component along q.  Returns 0."""
        fy, fc, fe = y.fields(), corr.fields(), err.fields()
        check(self.L.ampe_apply_projection(self.h, C.byref(fy), C.byref(fc), C.byref(fe), self._stream()),
              "applyProjection")
        return 0

    def computeSymmetryRotations(self, y):
        """QuatModel::computeSymmetryRotations (QuatModel.cc:4978-5055): rotation index of every
        lower face from y['quat'], kept in the context for the symmetry-aware evaluation"""
        fy = y.fields()
        check(self.L.ampe_rhs_compute_symmetry_rotations(self.h, C.byref(fy), self._stream()),
              "computeSymmetryRotations")

    def symmetryRotations(self):
        """copies of the context's rotation indices: one int32 tensor (ghost 0) per direction"""
        out = [torch.empty(self.ncell, dtype=torch.int32, device=self.device) for _ in range(self.cfg.ndim)]
        arr = (C.c_void_p * 3)()
        for d, t in enumerate(out):
            arr[d] = t.data_ptr()
        check(self.L.ampe_rhs_get_symmetry_rotations(self.h, arr, self._stream()), "get_rotations")
        return out

    def makeQuatFundamental(self, y):
        """QuatModel::makeQuatFundamental (QuatModel.cc:5059-5104), y['quat'] in place"""
        fy = y.fields()
        check(self.L.ampe_quat_fundamental(self.h, C.byref(fy), self._stream()), "makeQuatFundamental")

    def phaseConcentrations(self):
        """copies of the ctx-owned c_l, c_a (ghost-0) after an evaluation"""
        cl = torch.empty(self.ncell, dtype=torch.float64, device=self.device)
        ca = torch.empty(self.ncell, dtype=torch.float64, device=self.device)
        check(self.L.ampe_rhs_copy_phase_concentrations(self.h, cl.data_ptr(), ca.data_ptr(),
                                                        self._stream()), "copy cl/ca")
        return cl, ca

    def newtonFailures(self):
        return self.L.ampe_rhs_newton_failures(self.h, self._stream())

    def lastLaunchCount(self):
        return self.L.ampe_rhs_last_launch_count(self.h)

    def setKernelTiming(self, on=True):
        check(self.L.ampe_rhs_set_kernel_timing(self.h, 1 if on else 0), "setKernelTiming")

    def lastKernelMs(self):
        """(KKS pre-pass, fused kernel) device times of the last whole-slab evaluation, milliseconds"""
        a, b = C.c_double(0.0), C.c_double(0.0)
        check(self.L.ampe_rhs_last_kernel_ms(self.h, C.byref(a), C.byref(b)), "lastKernelMs")
        return a.value, b.value


def to_device(state, device="cuda"):
    """dict of CPU tensors (fields.make_state) -> SolutionVector on the GPU"""
    out = SolutionVector()
    for k in COMPONENTS:
        v = state.get(k)
        out[k] = None if v is None else v.to(device).contiguous()
    return out
