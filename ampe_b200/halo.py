"""Slab decomposition along the slowest axis and the ghost-plane exchange that
replaces SAMRAI's RefineSchedule::fillData on the RHS path
(reference: QuatIntegrator::fillScratch, source/QuatIntegrator.cc:2873-2955).

One process per GPU; rank r owns planes [r*ns, (r+1)*ns) of a periodic global
domain.  Per evaluation each rank sends its first/last `ng` planes of every
state component to its two neighbours (NCCL send/recv over NVLink, or gloo on
CPU for the tests) and evaluates the interior planes while the messages are in
flight; no collective is involved (the RHS has none: SURVEY.md 8e)."""
import os

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
    """Ghost planes of every state component, exchanged as ONE message per direction: the boundary
    planes of all components are gathered into a flat send buffer (one `torch.cat` kernel per
    direction), and the receive buffers are flat too, `lo[k]` / `hi[k]` being per-component views
    of them (component stride = ng planes, what ampe_rhs_set_halo expects)."""

    # ghost messages up to this size go through ONE all-gather of every rank's boundary planes
    # (a single NCCL call: the 2D workloads evaluate in ~0.1 ms, the host-side cost of posting
    # four point-to-point operations is of that order); larger ones are sent point to point
    ALLGATHER_MAX_BYTES = 1 << 20

    def __init__(self, ndim, ng, rank, nranks, group=None, mode=None):
        self.ndim, self.ng, self.rank, self.nranks, self.group = ndim, ng, rank, nranks, group
        self.prev = (rank - 1) % nranks
        self.next = (rank + 1) % nranks
        self.lo = self.hi = None
        self._send = None
        self.mode = mode or os.environ.get("AMPE_B200_HALO_MODE")  # None: by size; "p2p" | "allgather"

    @staticmethod
    def _with_depth(t):
        return t if t.dim() == 4 else t.unsqueeze(0)

    def _alloc(self, y):
        self.lo, self.hi = SolutionVector(), SolutionVector()
        present = [k for k in COMPONENTS if y.get(k) is not None]
        t0 = y[present[0]]
        shapes = {k: list(slab_planes(self._with_depth(y[k]), self.ndim, slice(0, self.ng)).shape)
                  for k in present}
        depth = sum(shapes[k][0] for k in present)
        flat = [depth] + shapes[present[0]][1:]
        mk = lambda: torch.empty(flat, dtype=t0.dtype, device=t0.device)
        if self.mode is None:
            nbytes = t0.element_size()
            for n in flat:
                nbytes *= n
            self.mode = "allgather" if nbytes <= self.ALLGATHER_MAX_BYTES else "p2p"
        if self.mode == "allgather":
            # [rank][0 low planes | 1 high planes][component planes ...]
            self._gather = torch.empty([self.nranks, 2] + flat, dtype=t0.dtype, device=t0.device)
            self._mine = torch.empty([2] + flat, dtype=t0.dtype, device=t0.device)
            self._send = (self._mine[0], self._mine[1])
            self._recv_lo = self._gather[self.prev, 1]   # lower neighbour's high planes
            self._recv_hi = self._gather[self.next, 0]   # upper neighbour's low planes
        else:
            self._send = (mk(), mk())       # my low planes (to prev), my high planes (to next)
            self._recv_lo, self._recv_hi = mk(), mk()
        off = 0
        for k in COMPONENTS:
            if k not in present:
                self.lo[k] = self.hi[k] = None
                continue
            d = shapes[k][0]
            out_shape = list(slab_planes(y[k], self.ndim, slice(0, self.ng)).shape)
            self.lo[k] = self._recv_lo[off:off + d].view(out_shape)
            self.hi[k] = self._recv_hi[off:off + d].view(out_shape)
            off += d
        self._present = present

    def start(self, y):
        """pack the boundary planes and post the sends / receives; returns the work handles"""
        if self.lo is None:
            self._alloc(y)
        ng = self.ng
        ts = [self._with_depth(y[k]) for k in self._present]
        n = ts[0].shape[slab_dim(self.ndim)]
        torch.cat([slab_planes(t, self.ndim, slice(0, ng)) for t in ts], 0, out=self._send[0])
        torch.cat([slab_planes(t, self.ndim, slice(n - ng, n)) for t in ts], 0, out=self._send[1])
        if self.mode == "allgather":
            return [dist.all_gather_into_tensor(self._gather.view(-1), self._mine.view(-1), group=self.group,
                                                async_op=True)]
        # order matters when prev == next (2 ranks): the peer's LOW planes are my HIGH ghosts
        ops = [dist.P2POp(dist.isend, self._send[0], self.prev, self.group),
               dist.P2POp(dist.isend, self._send[1], self.next, self.group),
               dist.P2POp(dist.irecv, self._recv_hi, self.next, self.group),
               dist.P2POp(dist.irecv, self._recv_lo, self.prev, self.group)]
        return dist.batch_isend_irecv(ops)

    @staticmethod
    def finish(works):
        for w in works:
            w.wait()

    def ghosted(self, t):
        """(depth, planes + 2 ng, ...) copy of a field with its neighbour planes (used once per
        time step for the Newton reference concentrations)"""
        y = SolutionVector({"phase": t, "quat": None, "conc": None, "temperature": None})
        tmp = SlabHalo(self.ndim, self.ng, self.rank, self.nranks, self.group, mode="p2p")
        tmp.finish(tmp.start(y))
        if t.is_cuda:
            torch.cuda.current_stream().synchronize()
        return torch.cat([tmp.lo["phase"], t, tmp.hi["phase"]], dim=slab_dim(self.ndim)).contiguous()


class DistributedRHS:
    """evaluateRHSFunction on a slab-decomposed periodic domain, halo exchange
    overlapped with the interior evaluation.

    The whole step (pack, NCCL send/recv on a communication stream, interior kernels, boundary
    kernels) is captured into a CUDA graph per (y, y_dot, fd_flag) buffer set and replayed:
    for the 2D workloads one evaluation is ~0.1 ms, less than the host-side cost of issuing
    the exchange.  `use_graphs=False` (or a failed capture) runs the same sequence eagerly."""

    def __init__(self, rhs, rank, nranks, group=None, use_graphs=None):
        self.rhs = rhs
        cfg = rhs.cfg
        self.halo = SlabHalo(cfg.ndim, rhs.nghosts(), rank, nranks, group)
        # the exchange must not queue behind the interior kernel's blocks: high-priority stream
        # (bench.py also sets TORCH_NCCL_HIGH_PRIORITY=1 for NCCL's own stream)
        self.comm_stream = torch.cuda.Stream(priority=-1)
        self.interior_first = os.environ.get("AMPE_B200_HALO_ORDER", "interior") == "interior"
        self._set = False
        # opt-in (AMPE_B200_GRAPHS=1): capturing NCCL send/recv needs a quiescent communicator on
        # every rank; the eager sequence is the default
        if use_graphs is None:
            use_graphs = os.environ.get("AMPE_B200_GRAPHS") == "1"
        self.use_graphs = bool(use_graphs)
        self._graphs = {}
        self._seen = {}
        self._launches = 0

    def resetRefPhaseConcentrations(self, cl_ref, ca_ref):
        ndim = self.rhs.cfg.ndim
        shp = (1, self.rhs.cfg.n[2] if ndim == 3 else 1, self.rhs.cfg.n[1], self.rhs.cfg.n[0])
        g0 = self.halo.ghosted(cl_ref.reshape(shp))
        g1 = self.halo.ghosted(ca_ref.reshape(shp))
        self.rhs.setRefPhaseConcentrationsGhosted(g0, g1)
        self._ref = (g0, g1)

    def _eager(self, time, y, y_dot, fd_flag):
        main = torch.cuda.current_stream()
        if not self._set:
            # first call: allocate the ghost buffers and hand their addresses to the context
            self.halo.finish(self.halo.start(y))
            self.rhs.setHalo(self.halo.lo, self.halo.hi)
            self._set = True
        # y is ready once `main` reaches this point; the interior planes are launched FIRST so
        # that the GPU computes while the host posts the exchange
        if self.interior_first:
            ready = main.record_event()
            self.rhs.evaluateRHSFunction(time, y, y_dot, fd_flag, part=1)  # interior planes
            self.comm_stream.wait_event(ready)
            with torch.cuda.stream(self.comm_stream):
                works = self.halo.start(y)
        else:
            self.comm_stream.wait_stream(main)
            with torch.cuda.stream(self.comm_stream):
                works = self.halo.start(y)
            self.rhs.evaluateRHSFunction(time, y, y_dot, fd_flag, part=1)  # interior planes
        with torch.cuda.stream(self.comm_stream):
            self.halo.finish(works)
        main.wait_stream(self.comm_stream)
        self.rhs.evaluateRHSFunction(time, y, y_dot, fd_flag, part=2)  # boundary planes
        self._launches = self.rhs.lastLaunchCount()

    def lastLaunchCount(self):
        return self._launches

    def evaluateRHSFunction(self, time, y, y_dot, fd_flag=0):
        if not self.use_graphs:
            self._eager(time, y, y_dot, fd_flag)
            return 0
        key = (tuple(0 if y.get(k) is None else y[k].data_ptr() for k in COMPONENTS),
               tuple(0 if y_dot.get(k) is None else y_dot[k].data_ptr() for k in COMPONENTS),
               int(fd_flag != 0))
        g = self._graphs.get(key)
        if g is not None:
            g.replay()
            return 0
        # the first two evaluations of a buffer set run eagerly (allocations, kernel attributes,
        # NCCL channel set-up all happen there); the third is captured
        seen = self._seen.get(key, 0)
        self._seen[key] = seen + 1
        if seen < 2 or len(self._graphs) >= 16:
            self._eager(time, y, y_dot, fd_flag)
            return 0
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._eager(time, y, y_dot, fd_flag)
            self._graphs[key] = g
            g.replay()
        except Exception as e:  # capture not possible here: stay eager
            import sys
            print("ampe_b200.halo: CUDA graph capture failed (%r); running eagerly" % (e,), file=sys.stderr)
            self.use_graphs = False
            torch.cuda.synchronize()
            self._eager(time, y, y_dot, fd_flag)
        return 0
