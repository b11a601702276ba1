"""N > 1 on the CPU (gloo, world_size 2): the per-rank scalar diagnostics of two slabs combined by
ampe_b200.diagnostics.combine_scalar_diagnostics equal the diagnostics of the whole domain (the MPI reductions
inside QuatModel::printScalarDiagnostics).  The per-rank numbers come from the CPU restatement here; on the GPU
they come from ampe_scalar_diagnostics."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path.insert(0, here)
        sys.path.insert(0, os.path.dirname(here))
        import numpy as np
        import parity
        from ampe_b200.diagnostics import combine_scalar_diagnostics
        from oracle import pyoracle
        cfg, st = parity.make_case(name)
        full = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
        o = pyoracle.Oracle(cfg)
        want = o.scalar_diagnostics(full)
        o.close()
        # this rank's slab along the slowest axis
        slab = cfg.ndim - 1
        npl = cfg.n[slab] // world
        assert npl * world == cfg.n[slab]
        cfg_r, _ = parity.make_case(name)
        cfg_r.n[slab] = npl
        cfg_r.nranks, cfg_r.rank = world, rank
        ax = -3 if cfg.ndim == 3 else -2
        sl = [slice(None)] * 4
        local = {}
        for k, v in full.items():
            if v is None:
                local[k] = None
                continue
            idx = [slice(None)] * v.ndim
            idx[v.ndim + ax] = slice(rank * npl, (rank + 1) * npl)
            local[k] = np.ascontiguousarray(v[tuple(idx)])
        o = pyoracle.Oracle(cfg_r)
        mine = o.scalar_diagnostics(local)
        o.close()
        got = combine_scalar_diagnostics(mine)
        for k, v in want.items():
            assert abs(got[k] - v) <= 1e-12 * max(1.0, abs(v)), (k, got[k], v)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["dendrite2d", "auni3d"])
def test_combined_diagnostics_equal_the_whole_domain(name):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_single_process_passthrough():
    from ampe_b200.diagnostics import combine_scalar_diagnostics
    d = {"volume": 2.0, "volume_solid": 1.0}
    assert combine_scalar_diagnostics(d) == d
