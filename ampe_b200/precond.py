"""Host-side handle of one block-preconditioner solver (include/ampe_b200_precond.h, csrc/mg.cu):
the device multigrid that replaces EllipticFACSolver / QuatSysSolver + hypre on the single periodic
level (reference: source/EllipticFACOps.h:35, QuatLevelSolver.cc:753-1080, QuatIntegrator.cc:3300-3771).
Arrays are torch CUDA float64 tensors in SAMRAI layout; torch only owns the memory."""
import ctypes as C

import torch

from .lib import AmpeError, check, load


def _ptrs(tensors):
    if tensors is None:
        return None
    arr = (C.c_void_p * 3)()
    for d, t in enumerate(tensors):
        if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
            raise AmpeError("side arrays must be contiguous CUDA float64 tensors")
        arr[d] = t.data_ptr()
    return arr


def _p(t):
    return None if t is None else t.data_ptr()


class LevelSolver:
    """ampe_mg: create / setOperatorCoefficients / solveSystem on the device"""

    def __init__(self, n=None, dx=None, with_column_scale=False, handle=None, owner=None, ncomp=1):
        self.L = load()
        self._owner = owner  # borrowed handle (a HostQuatIntegrator's block solver)
        self._keep = None
        if handle is not None:
            self.h, self._own = C.c_void_p(handle), False
            return
        ndim = len(n)
        nn = (C.c_int * 3)(*(list(n) + [1] * (3 - ndim)))
        hh = (C.c_double * 3)(*(list(dx) + [0.0] * (3 - ndim)))
        self.h = C.c_void_p()
        check(self.L.ampe_mg_create_multi(ndim, nn, hh, 1 if with_column_scale else 0, int(ncomp), C.byref(self.h)),
              "ampe_mg_create")
        self._own = True

    def close(self):
        if getattr(self, "_own", False) and self.h:
            self.L.ampe_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_elliptic(self, m=None, ngm=0, m_const=0.0, c=None, ngc=0, c_const=0.0, d=None, d2=None, ngd=0,
                     d_scale=1.0, d_const=0.0):
        """EllipticFACOps::setM / setC / setD*: M div(D grad u) + C u, D = d_scale * (d [+ d2]) or d_const"""
        self._keep = (m, c, d, d2)
        check(self.L.ampe_mg_set_elliptic(self.h, _p(m), ngm, m_const, _p(c), ngc, c_const, _ptrs(d), _ptrs(d2), ngd,
                                          d_scale, d_const, None), "ampe_mg_set_elliptic")

    def set_quat(self, gamma, mobility, ngm, face_coef, ngfc):
        """QuatFACOps::setOperatorCoefficients -> QuatLevelSolver::setMatrixCoefficients"""
        self._keep = (mobility, face_coef)
        check(self.L.ampe_mg_set_quat(self.h, gamma, _p(mobility), ngm, _ptrs(face_coef), ngfc, None),
              "ampe_mg_set_quat")

    def solve(self, rhs, ncycles=2, symmetrized=False, out=None, stream=None):
        """stream: a torch.cuda.Stream (None = the legacy default stream; AMPE_B200_MG_GRAPH=1 needs a real one)"""
        out = torch.empty_like(rhs) if out is None else out
        st = None if stream is None else C.c_void_p(stream.cuda_stream)
        check(self.L.ampe_mg_solve(self.h, rhs.data_ptr(), out.data_ptr(), int(ncycles), 1 if symmetrized else 0,
                                   st), "ampe_mg_solve")
        return out

    def apply(self, u):
        out = torch.empty_like(u)
        check(self.L.ampe_mg_apply(self.h, u.data_ptr(), out.data_ptr(), None), "ampe_mg_apply")
        return out

    def set_zero_slope(self, zero_slope):
        """homogeneous Neumann boundary per direction (before the set_* calls)"""
        zs = (C.c_int * 3)(*(list(zero_slope) + [0] * (3 - len(zero_slope))))
        check(self.L.ampe_mg_set_zero_slope(self.h, zs), "ampe_mg_set_zero_slope")

    def set_sweeps(self, pre, post, coarse):
        check(self.L.ampe_mg_set_sweeps(self.h, pre, post, coarse), "ampe_mg_set_sweeps")

    def num_levels(self):
        return self.L.ampe_mg_num_levels(self.h)

    def level_extents(self, level):
        n = (C.c_int * 3)()
        check(self.L.ampe_mg_level_extents(self.h, level, n), "ampe_mg_level_extents")
        return list(n)

    def level_array(self, level, which):
        n = self.level_extents(level)
        out = torch.empty((n[2], n[1], n[0]), dtype=torch.float64, device="cuda")
        check(self.L.ampe_mg_copy_level(self.h, level, which, out.data_ptr(), None), "ampe_mg_copy_level")
        return out

    def last_launch_count(self):
        return self.L.ampe_mg_last_launch_count(self.h)


def phasefacops_setc(n, phi, ngphi, m, ngm, gamma, well_scale, well_type, c, ngc):
    """PhaseFACOps::setCOnPatchPrivate on the box [0, n-1]"""
    L = load()
    ndim = len(n)
    lo = (C.c_int * 3)(0, 0, 0)
    hi = (C.c_int * 3)(*([v - 1 for v in n] + [0] * (3 - ndim)))
    check(L.ampe_k_phasefacops_setc(ndim, lo, hi, phi.data_ptr(), ngphi, m.data_ptr(), ngm, gamma, well_scale,
                                    well_type.encode(), c.data_ptr(), ngc, None), "ampe_k_phasefacops_setc")
