"""Grain diagnostics (Grains::findAndNumberGrains + computeGrainVolumes, source/Grains.cc:263-520, 647-697): the
restatement against an independent flood fill (breadth-first search over face neighbours) on small grids --
periodic wrap, zero-slope boundaries that do not connect, grains that touch, the threshold itself."""
import collections

import numpy as np
import pytest

from ampe_b200 import configs
from oracle import pyoracle


def bfs_grains(phase, thr, periodic, dv):
    shape = phase.shape   # (nz, ny, nx) or (ny, nx)
    nd = phase.ndim
    inside = phase >= thr
    seen = np.zeros(shape, dtype=bool)
    out = {}
    strides = [int(np.prod(shape[a + 1:])) for a in range(nd)]
    for start in np.argwhere(inside):
        start = tuple(int(v) for v in start)
        if seen[start]:
            continue
        q, cells, lowest = collections.deque([start]), 0, None
        seen[start] = True
        while q:
            c = q.popleft()
            cells += 1
            idx = sum(ci * si for ci, si in zip(c, strides))
            lowest = idx if lowest is None else min(lowest, idx)
            for a in range(nd):
                for s in (-1, 1):
                    v = list(c)
                    v[a] += s
                    if v[a] < 0 or v[a] >= shape[a]:
                        if not periodic[nd - 1 - a] or shape[a] == 1:
                            continue
                        v[a] %= shape[a]
                    v = tuple(v)
                    if inside[v] and not seen[v]:
                        seen[v] = True
                        q.append(v)
        out[lowest] = cells * dv
    return out


def blobs(shape, seed, nblob=6):
    rng = np.random.default_rng(seed)
    grid = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    phi = np.zeros(shape)
    for _ in range(nblob):
        c = [rng.uniform(0, n) for n in shape]
        r = rng.uniform(2.0, 0.25 * min(n for n in shape if n > 1))
        d2 = 0.0
        for g, ci, n in zip(grid, c, shape):
            d = np.abs(g - ci)
            d = np.minimum(d, n - d)   # periodic distance: blobs wrap around the box
            d2 = d2 + d * d
        phi = np.maximum(phi, 0.5 * (1.0 - np.tanh((np.sqrt(d2) - r) / 1.2)))
    return phi


@pytest.mark.parametrize("ndim,shape,slope", [(2, (40, 36), (0, 0)), (2, (40, 36), (1, 1)), (2, (33, 47), (1, 0)),
                                              (3, (14, 20, 24), (0, 0, 0)), (3, (14, 20, 24), (0, 1, 1))])
def test_grains_against_flood_fill(ndim, shape, slope):
    cfg = configs.dendrite2d(nx=shape[-1], ny=shape[-2]) if ndim == 2 else configs.gg3d_hbsm(nx=shape[2], ny=shape[1], nz=shape[0])
    for d in range(ndim):
        cfg.zero_slope[d] = slope[d]
    phase = blobs(shape, seed=3 + ndim + sum(slope))
    dv = float(np.prod([cfg.dx[d] for d in range(ndim)]))
    periodic = [not cfg.zero_slope[d] for d in range(ndim)]
    ref = bfs_grains(phase, 0.85, periodic, dv)
    o = pyoracle.Oracle(cfg)
    got, num = o.grain_volumes({"phase": phase}, 0.85, numbers=True)
    o.close()
    assert len(ref) >= 2
    assert sorted(got) == sorted(ref)
    for k in ref:
        assert got[k] == pytest.approx(ref[k], rel=1e-12)
    # every cell of a grain carries the grain's number, every other cell -1
    assert ((num >= 0) == (phase >= 0.85)).all()
    assert set(np.unique(num[num >= 0]).tolist()) == set(ref)


def test_threshold_is_inclusive_and_wrap_connects():
    cfg = configs.dendrite2d(nx=8, ny=6)
    phase = np.zeros((6, 8))
    phase[2, 0] = 0.85          # exactly the threshold: inside (Grains.cc:353 uses >=)
    phase[2, 7] = 0.9           # touches (2, 0) through the periodic x boundary
    phase[4, 3] = 0.849999      # just below: no grain
    o = pyoracle.Oracle(cfg)
    got = o.grain_volumes({"phase": phase}, 0.85)
    assert list(got) == [2 * 8 + 0] and got[16] == pytest.approx(2 * cfg.dx[0] * cfg.dx[1], rel=1e-14)
    o.close()
    cfg.zero_slope[0] = 1       # a physical boundary does not connect the two cells
    o = pyoracle.Oracle(cfg)
    got = o.grain_volumes({"phase": phase}, 0.85)
    assert sorted(got) == [16, 23]
    o.close()
