"""SURVEY.md 8f on the device: the quaternion symmetry pre-pass (QUAT_SYMM_ROTATION, QUAT_FUNDAMENTAL,
QuatModel::computeSymmetryRotations / makeQuatFundamental) and the CVODE projection hook
(PROJECT{2,3}D, QuatIntegrator::applyProjection), through the C ABI, against the oracle.
Rotation indices are integers: exact.  Doubles: the kernels keep the reference's operation order
(--fmad=false), the bar is 1e-14 relative."""
import ctypes as C

import numpy as np
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu


def _lib():
    from ampe_b200 import lib
    L = lib.load()
    vp, ci = C.c_void_p, C.c_int
    L.ampe_k_quat_symm_rotation.restype = ci
    L.ampe_k_quat_symm_rotation.argtypes = [ci, vp, vp, vp, ci, ci, C.POINTER(vp), ci, vp]
    L.ampe_k_quat_fundamental.restype = ci
    L.ampe_k_quat_fundamental.argtypes = [ci, vp, vp, vp, vp, vp, ci, vp]
    L.ampe_k_project.restype = ci
    L.ampe_k_project.argtypes = [ci, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    return L


def _ivec(v):
    return (C.c_int * 3)(*(list(v) + [0] * (3 - len(v))))


def _periodic_ghost(a, g, ndim):
    pad = [(0, 0)] + [(g, g)] * ndim
    return np.ascontiguousarray(np.pad(a, pad, mode="wrap"))


def _side_shape(n, a, g):
    ext = [n[d] + 2 * g + (1 if d == a else 0) for d in range(len(n))]
    return tuple(reversed(ext))


def _random_q(rng, qlen, n):
    shape = (qlen,) + tuple(reversed(n))
    if qlen == 1:
        return rng.uniform(-np.pi, np.pi, size=shape)
    q = rng.normal(size=shape)
    return q / np.sqrt((q * q).sum(0))[None]


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("ndim,qlen", [(2, 4), (3, 4), (2, 2), (3, 2), (2, 1)])
def test_k_quat_symm_rotation_matches_oracle(ndim, qlen):
    from oracle import pyoracle
    L = _lib()
    n = (37, 21) if ndim == 2 else (19, 11, 7)
    rng = np.random.default_rng(100 + 10 * ndim + qlen)
    qg = _periodic_ghost(_random_q(rng, qlen, n), 1, ndim)
    # previous indices: zeros, valid, negative (conjugate) and out-of-range values
    rot0 = [rng.integers(-52, 53, size=_side_shape(n, a, 1)).astype(np.int32) for a in range(ndim)]
    for r in rot0:
        r[rng.random(r.shape) < 0.3] = 0
    want = [r.copy() for r in rot0]
    pyoracle.quat_symm_rotation(n, qg, 1, qlen, want, 1)
    for pass_ in range(2):  # second pass: the indices found seed the search (in/out argument)
        got = [torch.as_tensor(r).cuda() for r in (rot0 if pass_ == 0 else want)]
        ptrs = (C.c_void_p * 3)(*[t.data_ptr() for t in got])
        qd = torch.as_tensor(qg).cuda()
        rc = L.ampe_k_quat_symm_rotation(ndim, _ivec([0] * ndim), _ivec([v - 1 for v in n]), qd.data_ptr(), 1,
                                         qlen, ptrs, 1, None)
        assert rc == 0, L.ampe_last_error()
        torch.cuda.synchronize()
        if pass_ == 1:
            pyoracle.quat_symm_rotation(n, qg, 1, qlen, want, 1)
        for a in range(ndim):
            assert np.array_equal(got[a].cpu().numpy(), want[a]), (pass_, a)


@pytest.mark.parametrize("ndim,qlen", [(2, 4), (3, 4), (2, 2), (2, 1)])
def test_k_quat_fundamental_matches_oracle(ndim, qlen):
    from oracle import pyoracle
    L = _lib()
    n = (33, 17) if ndim == 2 else (17, 9, 6)
    rng = np.random.default_rng(200 + 10 * ndim + qlen)
    qg = _periodic_ghost(_random_q(rng, qlen, n), 2, ndim)
    want = qg.copy()
    pyoracle.quat_fundamental(n, want, 2, qlen)
    qd = torch.as_tensor(qg).cuda()
    rc = L.ampe_k_quat_fundamental(ndim, _ivec([0] * ndim), _ivec([v - 1 for v in n]), qd.data_ptr(),
                                   _ivec([-2] * ndim), _ivec([v + 1 for v in n]), qlen, None)
    assert rc == 0, L.ampe_last_error()
    torch.cuda.synchronize()
    got = qd.cpu().numpy()
    assert _rel(got, want) <= 1e-14
    assert not np.array_equal(got, qg)  # something was rotated


@pytest.mark.parametrize("ndim,depth", [(2, 2), (2, 4), (3, 4), (3, 3)])
def test_k_project_matches_oracle(ndim, depth):
    from oracle import pyoracle
    L = _lib()
    n = (29, 18) if ndim == 2 else (14, 9, 7)
    rng = np.random.default_rng(300 + 10 * ndim + depth)
    shape = (depth,) + tuple(reversed(n))
    q = rng.normal(size=shape) * 0.3 + 0.5
    err = rng.normal(size=shape)
    qg = _periodic_ghost(q, 1, ndim)  # q with ghost width 1; corr and err ghost 0 (as CVODE's vectors)
    corr_o, err_o = np.zeros(shape), err.copy()
    pyoracle.project(n, depth, qg, 1, corr_o, 0, err_o, 0)
    qd, cd, ed = torch.as_tensor(qg).cuda(), torch.full(shape, float("nan"), dtype=torch.float64).cuda(), \
        torch.as_tensor(err).cuda()
    lo, hi = _ivec([0] * ndim), _ivec([v - 1 for v in n])
    glo, ghi = _ivec([-1] * ndim), _ivec(list(n))
    rc = L.ampe_k_project(ndim, lo, hi, depth, qd.data_ptr(), glo, ghi, cd.data_ptr(), lo, hi, ed.data_ptr(),
                          lo, hi, None)
    assert rc == 0, L.ampe_last_error()
    torch.cuda.synchronize()
    assert _rel(cd.cpu().numpy(), corr_o) <= 1e-14
    assert _rel(ed.cpu().numpy(), err_o) <= 1e-14
    assert torch.equal(qd.cpu(), torch.as_tensor(qg))  # q is an input


def _symmetric_equivalent_case(name):
    """small symmetry-aware case whose left half is multiplied by a cubic rotation: the fields are
    physically continuous but the stored quaternions jump, which is what the pre-pass resolves"""
    cfg, st = parity.make_case(name)
    assert cfg.symmetry_aware and cfg.qlen == 4
    from oracle import pyoracle
    t = np.zeros((48, 4))
    pyoracle.lib().oracle_qr_table4(t.ctypes.data)
    q = st["quat"].numpy().copy()
    half = q.shape[-1] // 2
    r = t[34]
    a = [q[m, ..., :half].copy() for m in range(4)]
    q[0, ..., :half] = a[0] * r[0] - a[1] * r[1] - a[2] * r[2] - a[3] * r[3]
    q[1, ..., :half] = a[0] * r[1] + a[1] * r[0] + a[2] * r[3] - a[3] * r[2]
    q[2, ..., :half] = a[0] * r[2] + a[2] * r[0] + a[3] * r[1] - a[1] * r[3]
    q[3, ..., :half] = a[0] * r[3] + a[3] * r[0] + a[1] * r[2] - a[2] * r[1]
    st["quat"] = torch.as_tensor(np.ascontiguousarray(q))
    return cfg, st


def _oracle_rotations(cfg, q):
    """QuatModel::computeSymmetryRotations with the oracle: lower faces of the interior cells"""
    from oracle import pyoracle
    ndim = cfg.ndim
    n = tuple(cfg.n[d] for d in range(ndim))
    qg = _periodic_ghost(q.reshape((cfg.qlen,) + tuple(reversed(n))), 1, ndim)  # 2D states carry nz = 1
    rot = [np.zeros(_side_shape(n, a, 1), dtype=np.int32) for a in range(ndim)]
    pyoracle.quat_symm_rotation(n, qg, 1, cfg.qlen, rot, 1)
    inner = tuple(slice(1, 1 + n[d]) for d in reversed(range(ndim)))
    return [np.ascontiguousarray(r[inner]).ravel() for r in rot]


def test_compute_symmetry_rotations_and_symmetry_aware_rhs():
    """rows a7/a10 fed by the device pre-pass: indices == oracle's, and the RHS evaluated with them
    == the oracle's RHS with the oracle's indices"""
    from ampe_b200 import rhs
    cfg, st = _symmetric_equivalent_case("auni2d")
    want = _oracle_rotations(cfg, st["quat"].numpy())
    assert any((w != 1).any() for w in want), "the case must need non-trivial rotations"
    y = rhs.to_device(st)
    r = rhs.QuatIntegratorRHS(cfg)
    r.computeSymmetryRotations(y)
    got = [t.cpu().numpy() for t in r.symmetryRotations()]
    for a in range(cfg.ndim):
        assert np.array_equal(got[a], want[a]), a
    # a second call starts from the indices found and must keep them
    r.computeSymmetryRotations(y)
    for a, t in enumerate(r.symmetryRotations()):
        assert np.array_equal(t.cpu().numpy(), want[a]), a
    c0 = y["conc"].reshape(-1).clone()
    r.resetRefPhaseConcentrations(c0, c0.clone())
    yd = y.like()
    r.evaluateRHSFunction(0.0, y, yd, 0)
    torch.cuda.synchronize()
    assert r.newtonFailures() == 0
    o_outs, _ = parity.run_oracle(cfg, st, (0,), want)
    status, yo = o_outs[0]
    assert status == 0
    ld, _ = parity.run_arbiter(cfg, st, (0,), want)
    for k in ("phase", "quat", "conc"):
        parity.check_one(k, yd[k].cpu().numpy(), yo[k], ld[0][k])
    r.close()


@pytest.mark.parametrize("name", ["auni2d", "dendrite2d", "auni3d"])
def test_make_quat_fundamental(name):
    from ampe_b200 import rhs
    from oracle import pyoracle
    cfg, st = parity.make_case(name)
    ndim = cfg.ndim
    n = tuple(cfg.n[d] for d in range(ndim))
    want = st["quat"].numpy().copy()
    pyoracle.quat_fundamental(n, want, 0, cfg.qlen)
    y = rhs.to_device(st)
    r = rhs.QuatIntegratorRHS(cfg)
    r.makeQuatFundamental(y)
    torch.cuda.synchronize()
    assert _rel(y["quat"].cpu().numpy(), want) <= 1e-14
    r.close()


@pytest.mark.parametrize("name", ["dendrite2d", "auni3d", "pfhub1a"])
def test_apply_projection(name):
    """QuatIntegrator::applyProjection: corr zeroed except the projected quaternion part"""
    from ampe_b200 import rhs
    from oracle import pyoracle
    cfg, st = parity.make_case(name)
    ndim = cfg.ndim
    n = tuple(cfg.n[d] for d in range(ndim))
    y = rhs.to_device(st)
    g = torch.Generator().manual_seed(5)
    if y.get("quat") is not None:  # an iterate that has drifted off the unit sphere
        y["quat"] = (y["quat"] * (1.0 + 0.05 * torch.rand(y["quat"].shape, generator=g, dtype=torch.float64).cuda())
                     ).contiguous()
    corr, err = y.like(), y.like()
    err_host = {}
    for k, v in err.items():
        if v is not None:
            e = torch.randn(v.shape, generator=g, dtype=torch.float64)
            err_host[k] = e.numpy().copy()
            v.copy_(e)
            corr[k].fill_(7.0)
    y_before = {k: (None if v is None else v.clone()) for k, v in y.items()}
    r = rhs.QuatIntegratorRHS(cfg)
    assert r.applyProjection(0.0, y, corr, 1e-10, err) == 0
    torch.cuda.synchronize()
    for k, v in y.items():
        if v is not None:
            assert torch.equal(v, y_before[k]), k  # y is an input
    evolved = {"phase": cfg.with_phase, "conc": cfg.with_concentration,
               "temperature": cfg.with_unsteady_temperature}
    for k, on in evolved.items():
        if on and corr.get(k) is not None:
            assert float(corr[k].abs().max()) == 0.0, k
            assert np.array_equal(err[k].cpu().numpy(), err_host[k]), k
    if cfg.evolve_quat and cfg.qlen > 1:
        q = y["quat"].cpu().numpy()
        corr_o, err_o = np.zeros_like(q), err_host["quat"].copy()
        pyoracle.project(n, cfg.qlen, q, 0, corr_o, 0, err_o, 0)
        assert _rel(corr["quat"].cpu().numpy(), corr_o) <= 1e-14
        assert _rel(err["quat"].cpu().numpy(), err_o) <= 1e-14
        qn = (y["quat"] + corr["quat"]).pow(2).sum(0)
        assert float((qn - 1.0).abs().max()) < 1e-14
    r.close()
