"""Development aid (no GPU): runs the LOGIC of tests/test_gpu_widening_zzz_precond.py on the CPU by
substituting stand-ins backed by the test infrastructure (oracle/pyoracle.py: HostMG, Oracle) for the device
objects (ampe_b200.precond.LevelSolver, ampe_b200.host_rhs.HostQuatIntegrator).  It proves nothing about the
CUDA code -- device-vs-CPU comparisons become CPU-vs-CPU -- but it catches shape, argument-order and
threshold mistakes in tests that cannot be executed until a GPU is available.
usage: python tools/check_gpu_test_logic.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["AMPE_B200_RUN_EXPERIMENTS"] = "1"

from oracle import pyoracle  # noqa: E402
import ampe_b200.host_rhs as host_rhs  # noqa: E402
import ampe_b200.precond as precond  # noqa: E402
import ampe_b200.rhs as rhs  # noqa: E402


class FakeLevelSolver:
    def __init__(self, n=None, dx=None, with_column_scale=False, handle=None, owner=None, ncomp=1):
        self.g = handle if handle is not None else pyoracle.HostMG(n, dx, with_s=with_column_scale, ncomp=ncomp)
        self.launches = 3 if os.environ.get("AMPE_B200_MG_TAIL") == "0" else 1  # the switch is read at creation

    @staticmethod
    def _np(t):
        return None if t is None else t.numpy()

    def set_elliptic(self, m=None, ngm=0, m_const=0.0, c=None, ngc=0, c_const=0.0, d=None, d2=None, ngd=0,
                     d_scale=1.0, d_const=0.0):
        ls = lambda x: None if x is None else [t.numpy() for t in x]
        self.g.set_elliptic(m=self._np(m), ngm=ngm, m_const=m_const, c=self._np(c), ngc=ngc, c_const=c_const,
                            d=ls(d), d2=ls(d2), ngd=ngd, d_scale=d_scale, d_const=d_const)

    def set_quat(self, gamma, mob, ngm, fc, ngfc):
        self.g.set_quat(gamma, mob.numpy(), ngm, [t.numpy() for t in fc], ngfc)

    def solve(self, rhs_, ncycles=2, symmetrized=False, out=None, stream=None):
        z = torch.as_tensor(self.g.solve(rhs_.numpy(), ncycles, symmetrized))
        if out is not None:
            out.copy_(z)
            return out
        return z

    def apply(self, u):
        return torch.as_tensor(self.g.apply(u.numpy()))

    def num_levels(self):
        return self.g.num_levels()

    def level_extents(self, level):
        return self.g.level_extents(level)

    def level_array(self, level, which):
        return torch.as_tensor(self.g.level_array(level, which))

    def last_launch_count(self):
        return self.launches

    def close(self):
        pass


def fake_setc(n, phi, ngphi, m, ngm, gamma, ws, wt, c, ngc):
    assert wt.startswith("d")
    nd = len(n)

    def interior(a, ng):
        a = a.numpy()
        if ng == 0:
            return a
        sl = tuple(slice(ng, -ng) if (3 - 1 - ax) < nd else slice(None) for ax in range(3))
        return a[sl]
    p, mm = interior(phi, ngphi), interior(m, ngm)
    c.copy_(torch.as_tensor(1.0 + (gamma * mm) * ws * (32.0 * (1.0 + 6.0 * p * (p - 1.0)))))


class FakeHost:
    IMPLICIT_ENEWTON = -20

    def __init__(self, cfg, use_fused):
        self.cfg, self.o, self.nc, self.dq, self.stats = cfg, pyoracle.Oracle(cfg), 0, False, [0.0, 0.0]

    @staticmethod
    def _np(y):
        return {k: (None if v is None else v.numpy()) for k, v in y.items()}

    def resetRefPhaseConcentrations(self, cl=None, ca=None):
        self.o.set_ref(cl.numpy().copy(), ca.numpy().copy())

    def setSymmetryRotations(self, iq):
        self.o.set_rotations([t.numpy() for t in iq])

    def evaluateRHSFunction(self, t, y, yd, fd=0):
        st, _ = self.o.eval(t, self._np(y), fd_flag=fd, ydot=self._np(yd))
        assert st == 0

    def setupPreconditioners(self, ncycles=2, precond_has_dquatdphi=False):
        self.nc, self.dq = ncycles, precond_has_dquatdphi
        self.o.set_preconditioner(ncycles, dquatdphi=precond_has_dquatdphi)

    def CVSpgmrPrecondSet(self, t, y, gamma):
        assert self.o.precond_setup(gamma, self.nc, dquatdphi=self.dq) == 0
        self.stats[0] += 1

    def CVSpgmrPrecondSolve(self, r, z):
        rc, _ = self.o.precond_solve(self._np(r), self._np(z))
        assert rc == 0
        self.stats[1] += 1

    def preconditionerLevelSolver(self, block):
        g = self.o.precond_block(block)
        return FakeLevelSolver(handle=g) if g is not None else None

    def multiplyDQuatDPhiBlock(self, phase, qlen):
        return torch.as_tensor(self.o.precond_dquatdphi(phase.numpy()))

    def precondStats(self):
        return {"precond_setups": self.stats[0], "precond_solves": self.stats[1]}

    def integrateImplicit(self, y, dt, nsteps, **kw):
        rc, st = self.o.integrate_implicit(self._np(y), dt, nsteps, **kw)
        self.stats = list(self.o.precond_stats().values())
        return rc, st

    def integrateAdaptive(self, y, tend, h0, **kw):
        return self.o.integrate_adaptive(self._np(y), tend, h0, **kw)

    def close(self):
        self.o.close()


class FakeRHS:
    def __init__(self, cfg, device=None):
        self.o = pyoracle.Oracle(cfg)

    def printScalarDiagnostics(self, y):
        return self.o.scalar_diagnostics({k: (None if v is None else v.numpy()) for k, v in y.items()})

    def close(self):
        self.o.close()


def main():
    rhs.QuatIntegratorRHS = FakeRHS
    precond.LevelSolver = FakeLevelSolver
    precond.phasefacops_setc = fake_setc
    host_rhs.HostQuatIntegrator = FakeHost
    real_to_device = rhs.to_device
    def to_host_copy(state, device="cpu"):  # a device copy never aliases the caller's CPU state
        out = real_to_device(state, "cpu")
        for k in list(out):
            out[k] = None if out[k] is None else out[k].clone()
        return out
    rhs.to_device = to_host_copy
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    real_empty, real_zeros = torch.empty, torch.zeros
    torch.empty = lambda *a, **k: real_empty(*a, **{kk: v for kk, v in k.items() if kk != "device"})
    torch.zeros = lambda *a, **k: real_zeros(*a, **{kk: v for kk, v in k.items() if kk != "device"})

    class FakeStream:
        def wait_stream(self, s):
            pass

        def synchronize(self):
            pass
    torch.cuda.Stream = FakeStream
    torch.cuda.current_stream = lambda *a, **k: None
    import test_gpu_widening_zzz_precond as T
    T._cuda = lambda a: torch.as_tensor(np.ascontiguousarray(a))
    ran = 0
    for name in sorted(dir(T)):
        fn = getattr(T, name)
        if not name.startswith("test_") or not callable(fn):
            continue
        marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
        cases = [()]
        if marks:
            cases = [c if isinstance(c, tuple) else (c,) for c in marks[0].args[1]]
        for c in cases:
            if "tmp_path" in fn.__code__.co_varnames[:fn.__code__.co_argcount]:
                import pathlib
                import tempfile
                with tempfile.TemporaryDirectory() as d:
                    fn(pathlib.Path(d))
                ran += 1
                print("ok", name, "(tmp_path)")
                continue
            fn(*c)
            ran += 1
            print("ok", name, c)
    print("%d test invocations exercised on the CPU stand-ins" % ran)


if __name__ == "__main__":
    main()
