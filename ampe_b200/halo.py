"""Slab decomposition along the slowest axis and the ghost-plane exchange that
replaces SAMRAI's RefineSchedule::fillData on the RHS path
(reference: QuatIntegrator::fillScratch, source/QuatIntegrator.cc:2873-2955).

One process per GPU; rank r owns planes [r*ns, (r+1)*ns) of a periodic global
domain.  The exchange itself lives behind the C ABI (csrc/halo.cu, ampe_halo_* /
ampe_rhs_eval_slab): every rank pushes its first/last `ng` planes of every state
component straight into its neighbours' HBM over NVLink (peer-mapped receive
buffers, epoch flags) and evaluates the interior planes meanwhile; `DistributedRHS`
only ships the opaque set-up handles between the ranks (torch.distributed stands
in for the MPI_Sendrecv AMPE would use).  `SlabHalo` is the torch.distributed
(NCCL send/recv, gloo on CPU) exchange of the same planes: transport "nccl" of
`DistributedRHS`, and what the gloo tests of the slab indexing run on."""
import ctypes as C
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


_HANDLE_BYTES = 128  # AMPE_HALO_HANDLE_BYTES (include/ampe_b200.h)


class DistributedRHS:
    """evaluateRHSFunction on a slab-decomposed periodic domain.

    transport "ipc" (default on a GPU): the C-ABI exchange, one call per evaluation (ampe_rhs_eval_slab).
    transport "nccl" (AMPE_B200_HALO=nccl): ghost planes by torch.distributed into buffers handed to
    ampe_rhs_set_halo, interior / boundary split driven from here."""

    def __init__(self, rhs, rank, nranks, group=None, transport=None):
        self.rhs = rhs
        self.rank, self.nranks, self.group = rank, nranks, group
        cfg = rhs.cfg
        self.transport = transport or os.environ.get("AMPE_B200_HALO", "ipc")
        self._launches = 0
        self.h = None
        if self.transport == "ipc":
            try:
                self._connect()
            except Exception as e:  # no CUDA IPC between these processes: the NCCL transport does the same exchange
                import sys
                print("ampe_b200.halo: peer-mapped exchange unavailable (%r); using the NCCL transport" % (e,),
                      file=sys.stderr)
                self.transport = "nccl"
        if self.transport != "ipc":
            self.halo = SlabHalo(cfg.ndim, rhs.nghosts(), rank, nranks, group)
            # the exchange must not queue behind the interior kernel's blocks: high-priority stream
            self.comm_stream = torch.cuda.Stream(priority=-1)
            self._set = False

    # ---- set-up of the C-ABI exchange: create, export, ship the handles, connect -----------------
    def _connect(self):
        from .lib import check
        L = self.rhs.L
        h = C.c_void_p()
        check(L.ampe_halo_create(self.rhs.h, self.rank, self.nranks, C.byref(h)), "ampe_halo_create")
        mine = C.create_string_buffer(_HANDLE_BYTES)
        check(L.ampe_halo_export(h, mine), "ampe_halo_export")
        blobs = [None] * self.nranks
        dist.all_gather_object(blobs, bytes(mine.raw), group=self.group)
        prev = C.create_string_buffer(blobs[(self.rank - 1) % self.nranks], _HANDLE_BYTES)
        nxt = C.create_string_buffer(blobs[(self.rank + 1) % self.nranks], _HANDLE_BYTES)
        rc = L.ampe_halo_connect(h, prev, nxt)
        # every rank must know whether every rank connected (a half-connected ring would hang in the first wait)
        ok = [None] * self.nranks
        dist.all_gather_object(ok, int(rc), group=self.group)
        if any(v != 0 for v in ok):
            L.ampe_halo_destroy(h)
            check(rc, "ampe_halo_connect")
            raise RuntimeError("a neighbour could not map the receive buffers (codes %r)" % (ok,))
        self.h = h
        dist.barrier(group=self.group)

    def close(self):
        if self.h is not None:
            self.rhs.L.ampe_halo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def resetRefPhaseConcentrations(self, cl_ref=None, ca_ref=None):
        """interior arrays (ghost 0) of this rank, or nothing = "the last computed c_l, c_a" (ghost planes
        included: QuatModel::resetRefPhaseConcentrations copies whole arrays)"""
        if cl_ref is None:
            self.rhs.resetRefPhaseConcentrations()
            return
        if self.transport == "ipc":
            from .lib import check
            self._keep_ref = (cl_ref, ca_ref)
            check(self.rhs.L.ampe_rhs_set_ref_concentrations_slab(self.rhs.h, self.h, cl_ref.data_ptr(),
                                                                  ca_ref.data_ptr(), self._stream()), "set_ref_slab")
            return
        ndim = self.rhs.cfg.ndim
        shp = (1, self.rhs.cfg.n[2] if ndim == 3 else 1, self.rhs.cfg.n[1], self.rhs.cfg.n[0])
        g0 = self.halo.ghosted(cl_ref.reshape(shp))
        g1 = self.halo.ghosted(ca_ref.reshape(shp))
        self.rhs.setRefPhaseConcentrationsGhosted(g0, g1)
        self._ref = (g0, g1)

    def setSymmetryRotations(self, iqrot):
        """rotation indices of this rank's lower faces (ghost 0); the ghost planes come from the neighbours"""
        if self.transport != "ipc":
            raise RuntimeError("the symmetry-aware path on several ranks needs the peer-mapped exchange")
        from .lib import check
        arr = (C.c_void_p * 3)()
        self._iq = [t.to(torch.int32).contiguous() for t in iqrot]
        for d, t in enumerate(self._iq):
            arr[d] = t.data_ptr()
        check(self.rhs.L.ampe_rhs_set_symmetry_rotations_slab(self.rhs.h, self.h, arr, self._stream()),
              "set_rotations_slab")

    def integrateFixed(self, y, dt, nsteps, scheme=0, t0=0.0):
        """nsteps explicit steps (0 Euler, 1 Heun) of this rank's slab on the device, ghost planes exchanged at every
        evaluation; y updated in place (QuatIntegratorRHS.integrateFixed on several ranks)"""
        from .lib import check
        w1 = y.like()
        w2 = y.like() if scheme == 1 else None
        fy, f1 = y.fields(), w1.fields()
        f2 = w2.fields() if w2 is not None else None
        check(self.rhs.L.ampe_integrate_fixed_slab(self.rhs.h, self.h, C.byref(fy), C.byref(f1),
                                                   C.byref(f2) if f2 is not None else None, float(t0), float(dt),
                                                   int(nsteps), int(scheme), self._stream()), "integrateFixed (slab)")
        torch.cuda.current_stream().synchronize()

    def computeSymmetryRotations(self, y):
        """QuatModel::computeSymmetryRotations on this rank's slab (ghost planes of y and of the indices from the
        neighbours)"""
        from .lib import check
        fy = y.fields()
        check(self.rhs.L.ampe_rhs_compute_symmetry_rotations_slab(self.rhs.h, self.h, C.byref(fy), self._stream()),
              "computeSymmetryRotations (slab)")

    def _eager(self, time, y, y_dot, fd_flag):
        main = torch.cuda.current_stream()
        if not self._set:
            # first call: allocate the ghost buffers and hand their addresses to the context
            self.halo.finish(self.halo.start(y))
            self.rhs.setHalo(self.halo.lo, self.halo.hi)
            self._set = True
        # y is ready once `main` reaches this point; the interior planes are launched FIRST so
        # that the GPU computes while the host posts the exchange
        ready = main.record_event()
        self.rhs.evaluateRHSFunction(time, y, y_dot, fd_flag, part=1)  # interior planes
        self.comm_stream.wait_event(ready)
        with torch.cuda.stream(self.comm_stream):
            works = self.halo.start(y)
            self.halo.finish(works)
        main.wait_stream(self.comm_stream)
        self.rhs.evaluateRHSFunction(time, y, y_dot, fd_flag, part=2)  # boundary planes
        self._launches = self.rhs.lastLaunchCount() + 2

    def lastLaunchCount(self):
        return self._launches

    def evaluateRHSFunction(self, time, y, y_dot, fd_flag=0):
        if self.transport != "ipc":
            self._eager(time, y, y_dot, fd_flag)
            return 0
        from .lib import check
        fy, fd = y.fields(), y_dot.fields()
        check(self.rhs.L.ampe_rhs_eval_slab(self.rhs.h, self.h, float(time), C.byref(fy), C.byref(fd), int(fd_flag),
                                            self._stream()), "ampe_rhs_eval_slab")
        self._launches = self.rhs.L.ampe_halo_last_launch_count(self.h)
        return 0

    def evaluateRHSFunctionHost(self, time, y_host, ydot_host, fd_flag=0):
        """HOST buffers of this rank's slab (pinned): chunk pipeline H2D | kernels | D2H, ghost planes device to
        device"""
        if self.transport != "ipc":
            raise RuntimeError("the host-buffer path on several ranks needs the peer-mapped exchange")
        from . import _abi
        from .lib import check
        fy, fd = _abi.RhsFields(), _abi.RhsFields()
        for k in COMPONENTS:
            a, b = y_host.get(k), ydot_host.get(k)
            setattr(fy, k, None if a is None else a.data_ptr())
            setattr(fd, k, None if b is None else b.data_ptr())
        check(self.rhs.L.ampe_rhs_eval_slab_host(self.rhs.h, self.h, float(time), C.byref(fy), C.byref(fd),
                                                 int(fd_flag)), "ampe_rhs_eval_slab_host")
        return 0
