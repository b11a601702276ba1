"""Slab decomposition along the slowest axis and the ghost-plane exchange that
replaces SAMRAI's RefineSchedule::fillData on the RHS path
(reference: QuatIntegrator::fillScratch, source/QuatIntegrator.cc:2873-2955).

One process per GPU; rank r owns planes [r*ns, (r+1)*ns) of a periodic global
domain.  Per evaluation each rank sends its first/last `ng` planes of every
state component to its two neighbours (NCCL send/recv over NVLink, or gloo on
CPU for the tests) and evaluates the interior planes while the messages are in
flight; no collective is involved (the RHS has none: SURVEY.md 8e)."""
import torch
import torch.distributed as dist

from .rhs import COMPONENTS, SolutionVector


def slab_dim(ndim):
    """the slab axis counted from the end: z of (..., nz, ny, nx) in 3D, y of (..., ny, nx) in 2D
    (components may or may not carry a leading depth dimension)"""
    return -3 if ndim == 3 else -2


def slab_planes(t, ndim, sl):
    """view of planes `sl` along the slab axis of a (..., nz, ny, nx) tensor"""
    return t[..., sl, :, :] if ndim == 3 else t[..., sl, :]


class SlabHalo:
    def __init__(self, ndim, ng, rank, nranks, group=None):
        self.ndim, self.ng, self.rank, self.nranks, self.group = ndim, ng, rank, nranks, group
        self.prev = (rank - 1) % nranks
        self.next = (rank + 1) % nranks
        self.lo = self.hi = None
        self._send = None

    def _alloc(self, y):
        self.lo, self.hi, self._send = SolutionVector(), SolutionVector(), {}
        for k in COMPONENTS:
            t = y.get(k)
            if t is None:
                self.lo[k] = self.hi[k] = None
                continue
            shape = list(slab_planes(t, self.ndim, slice(0, self.ng)).shape)
            self.lo[k] = torch.empty(shape, dtype=t.dtype, device=t.device)
            self.hi[k] = torch.empty(shape, dtype=t.dtype, device=t.device)
            self._send[k] = (torch.empty(shape, dtype=t.dtype, device=t.device),
                             torch.empty(shape, dtype=t.dtype, device=t.device))

    def start(self, y):
        """pack the boundary planes and post the sends / receives; returns the work handles"""
        if self.lo is None:
            self._alloc(y)
        ng = self.ng
        ops = []
        recv = []
        for k in COMPONENTS:
            t = y.get(k)
            if t is None:
                continue
            s_low, s_high = self._send[k]
            s_low.copy_(slab_planes(t, self.ndim, slice(0, ng)))
            s_high.copy_(slab_planes(t, self.ndim, slice(t.shape[slab_dim(self.ndim)] - ng, None)))
            ops.append(dist.P2POp(dist.isend, s_low, self.prev, self.group))
            ops.append(dist.P2POp(dist.isend, s_high, self.next, self.group))
            # order matters when prev == next (2 ranks): the peer's LOW planes are my HIGH ghosts
            recv.append(dist.P2POp(dist.irecv, self.hi[k], self.next, self.group))
            recv.append(dist.P2POp(dist.irecv, self.lo[k], self.prev, self.group))
        return dist.batch_isend_irecv(ops + recv)

    @staticmethod
    def finish(works):
        for w in works:
            w.wait()

    def ghosted(self, t):
        """(depth, planes + 2 ng, ...) copy of a field with its neighbour planes (used once per
        time step for the Newton reference concentrations)"""
        y = SolutionVector({"phase": t, "quat": None, "conc": None, "temperature": None})
        tmp = SlabHalo(self.ndim, self.ng, self.rank, self.nranks, self.group)
        tmp.finish(tmp.start(y))
        if t.is_cuda:
            torch.cuda.current_stream().synchronize()
        return torch.cat([tmp.lo["phase"], t, tmp.hi["phase"]], dim=slab_dim(self.ndim)).contiguous()


class DistributedRHS:
    """evaluateRHSFunction on a slab-decomposed periodic domain, halo exchange
    overlapped with the interior evaluation."""

    def __init__(self, rhs, rank, nranks, group=None):
        self.rhs = rhs
        cfg = rhs.cfg
        self.halo = SlabHalo(cfg.ndim, rhs.nghosts(), rank, nranks, group)
        self.comm_stream = torch.cuda.Stream()
        self._set = False

    def resetRefPhaseConcentrations(self, cl_ref, ca_ref):
        ndim = self.rhs.cfg.ndim
        shp = (1, self.rhs.cfg.n[2] if ndim == 3 else 1, self.rhs.cfg.n[1], self.rhs.cfg.n[0])
        g0 = self.halo.ghosted(cl_ref.reshape(shp))
        g1 = self.halo.ghosted(ca_ref.reshape(shp))
        self.rhs.setRefPhaseConcentrationsGhosted(g0, g1)
        self._ref = (g0, g1)

    def evaluateRHSFunction(self, time, y, y_dot, fd_flag=0):
        main = torch.cuda.current_stream()
        self.comm_stream.wait_stream(main)
        with torch.cuda.stream(self.comm_stream):
            works = self.halo.start(y)
        if not self._set:
            self.rhs.setHalo(self.halo.lo, self.halo.hi)
            self._set = True
        self.rhs.evaluateRHSFunction(time, y, y_dot, fd_flag, part=1)  # interior planes
        with torch.cuda.stream(self.comm_stream):
            self.halo.finish(works)
        main.wait_stream(self.comm_stream)
        self.rhs.evaluateRHSFunction(time, y, y_dot, fd_flag, part=2)  # boundary planes
        return 0
