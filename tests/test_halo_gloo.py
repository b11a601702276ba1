"""Host-side logic of the slab decomposition (ampe_b200/halo.py) on CPU with the
gloo backend, world_size 2 and 3: after the exchange every rank's ghost planes
equal the periodic neighbour planes of the global field (what SAMRAI's
RefineSchedule::fillData provides in the reference, QuatIntegrator.cc:2948-2954)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ndim, ng, q, mode=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ampe_b200.halo import SlabHalo, slab_planes
        from ampe_b200.rhs import SolutionVector
        nx, ny, nzl = 6, 5, 4
        if ndim == 3:
            gshape = (nzl * world, ny, nx)
        else:
            gshape = (1, nzl * world, nx)
        g = torch.Generator().manual_seed(1234)
        glob = {"phase": torch.rand((1,) + gshape, generator=g, dtype=torch.float64),
                "quat": torch.rand((4,) + gshape, generator=g, dtype=torch.float64),
                # a component WITHOUT the leading depth dimension, like fields.make_state's phase
                "conc": torch.rand(gshape, generator=g, dtype=torch.float64),
                "temperature": torch.rand((1,) + gshape, generator=g, dtype=torch.float64)}
        dim = -3 if ndim == 3 else -2
        sl = slice(rank * nzl, (rank + 1) * nzl)
        y = SolutionVector({k: (None if v is None else slab_planes(v, ndim, sl).contiguous())
                            for k, v in glob.items()})
        h = SlabHalo(ndim, ng, rank, world, mode=mode)
        h.finish(h.start(y))
        ntot = nzl * world
        for k, v in glob.items():
            if v is None:
                assert h.lo[k] is None
                continue
            lo_idx = [(rank * nzl - ng + i) % ntot for i in range(ng)]
            hi_idx = [((rank + 1) * nzl + i) % ntot for i in range(ng)]
            exp_lo = v.index_select(dim, torch.tensor(lo_idx))
            exp_hi = v.index_select(dim, torch.tensor(hi_idx))
            assert torch.equal(h.lo[k], exp_lo), (rank, k, "lo")
            assert torch.equal(h.hi[k], exp_hi), (rank, k, "hi")
        # second exchange reuses the buffers
        y["phase"] += 1.0
        h.finish(h.start(y))
        assert torch.equal(h.lo["phase"], glob["phase"].index_select(
            dim, torch.tensor([(rank * nzl - ng + i) % ntot for i in range(ng)])) + 1.0)
        gh = h.ghosted(y["temperature"])
        assert gh.shape[dim] == nzl + 2 * ng
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ndim,ng,mode", [(2, 2, 1, "p2p"), (2, 3, 1, "p2p"), (3, 3, 1, "p2p"),
                                                (2, 2, 2, "p2p"), (2, 2, 1, "allgather"),
                                                (3, 3, 1, "allgather"), (2, 2, 2, None)])
def test_slab_halo_exchange(world, ndim, ng, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ndim, ng, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
