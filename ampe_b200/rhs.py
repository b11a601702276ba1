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


def to_device(state, device="cuda"):
    """dict of CPU tensors (fields.make_state) -> SolutionVector on the GPU"""
    out = SolutionVector()
    for k in COMPONENTS:
        v = state.get(k)
        out[k] = None if v is None else v.to(device).contiguous()
    return out
