"""Pin the CPU oracle against the reference's own known-answer tests:

  tests/CALPHADbinaryEquilibrium/test.input + tests/testCALPHADbinaryEquilibrium.cc:111-127  (golden)
  tests/testCALPHADbinaryKKS.cc:123-161                                                 (property)
  tests/testCALPHADFunctions.cc                                  (analytic vs finite difference)
  tests/testGradQ.cc:62-176                                         (linear field -> exact slopes)
  tests/testFlux.cc                                              (flux = D * slope / dx pattern)
  tests/testInterpolationFunctions.cc:23-63                                   (pbg identities)

These run on the CPU (no GPU needed)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from ampe_b200 import _abi, configs
from oracle import pyoracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))


@pytest.fixture(scope="module")
def L():
    return pyoracle.lib()


@pytest.fixture(scope="module")
def db():
    return configs.load_calphad()


def test_abi_struct_size(L):
    assert L.oracle_abi_sizeof_config() == C.sizeof(_abi.RhsConfig)


def test_calphad_equilibrium_golden(L, db):
    g = GOLD["CALPHADbinaryEquilibrium"]
    x = np.array(g["initial_guess"], dtype=np.float64)
    it = L.oracle_calphad_ceq(C.byref(db), g["temperature"], x.ctypes.data, 1e-8, 50, 1.0)
    assert it >= 0
    assert 0.0 <= x[0] <= 1.0 and 0.0 <= x[1] <= 1.0
    # the reference's (one-sided) check, tol 1e-6 -- and the two-sided one
    for k in range(2):
        assert (g["expected"][k] - x[k]) <= g["tol"]
        assert abs(g["expected"][k] - x[k]) <= g["tol"]


def test_calphad_kks_property(L, db):
    g = GOLD["CALPHADbinaryKKS"]
    hphi = L.oracle_interp_func(g["phi"], b"p")
    x = np.array([g["c"], g["c"]], dtype=np.float64)
    it = L.oracle_calphad_phase_concentrations(C.byref(db), g["temperature"], g["c"], hphi,
                                               x.ctypes.data, 1e-8, 50, 1.0)
    assert it >= 0
    mul = L.oracle_calphad_deriv_free_energy(C.byref(db), g["temperature"], x[0], 0)
    mus = L.oracle_calphad_deriv_free_energy(C.byref(db), g["temperature"], x[1], 1)
    assert abs(mul - mus) < g["tol"]
    assert abs((1 - hphi) * x[0] + hphi * x[1] - g["c"]) < 1e-8


def test_calphad_functions_fd(L, db):
    """testCALPHADFunctions.cc: analytic derivatives vs finite differences, eps 1e-8, tol 1e-6"""
    eps, tol = 1e-8, 1e-6
    l0, l1, l2, l3 = 2.3, 5.1, 3.2, -2.5
    for c in (0.33, 0.1, 0.78):
        f0 = L.oracle_calphad_fmix(l0, l1, l2, l3, c)
        f1 = L.oracle_calphad_fmix(l0, l1, l2, l3, c + eps)
        d = L.oracle_calphad_fmix_deriv(l0, l1, l2, l3, c)
        assert abs((f1 - f0) / eps - d) < tol
        # second derivative: central difference (the one-sided 1e-8 quotient of the reference
        # test carries truncation + 1e-16/eps rounding errors, both about 1e-6 here)
        h = 1e-5
        dm = L.oracle_calphad_fmix_deriv(l0, l1, l2, l3, c - h)
        dp = L.oracle_calphad_fmix_deriv(l0, l1, l2, l3, c + h)
        d2 = L.oracle_calphad_fmix_deriv2(l0, l1, l2, l3, c)
        assert abs((dp - dm) / (2 * h) - d2) < tol * max(1.0, abs(d2))
        assert abs((L.oracle_xlogx(c + eps) - L.oracle_xlogx(c)) / eps - L.oracle_xlogx_deriv(c)) < tol
        assert abs((L.oracle_xlogx_deriv(c + eps) - L.oracle_xlogx_deriv(c)) / eps -
                   L.oracle_xlogx_deriv2(c)) < 1e-5
    # full free energy: mu = df/dc and d2f = dmu/dc for both phases (relative: J/mol scale 1e4)
    T = 1450.0
    for pi in (0, 1):
        for c in (0.1, 0.25, 0.6):
            f0 = L.oracle_calphad_free_energy(C.byref(db), T, c - 1e-6, pi)
            f1 = L.oracle_calphad_free_energy(C.byref(db), T, c + 1e-6, pi)
            mu = L.oracle_calphad_deriv_free_energy(C.byref(db), T, c, pi)
            assert abs((f1 - f0) / 2e-6 - mu) < 1e-3 * max(1.0, abs(mu))
            m0 = L.oracle_calphad_deriv_free_energy(C.byref(db), T, c - 1e-6, pi)
            m1 = L.oracle_calphad_deriv_free_energy(C.byref(db), T, c + 1e-6, pi)
            d2 = L.oracle_calphad_second_deriv_free_energy(C.byref(db), T, c, pi)
            assert abs((m1 - m0) / 2e-6 - d2) < 1e-5 * abs(d2)
    # xlogx extension is C1 at the switch point
    s = 1e-8
    assert abs(L.oracle_xlogx(s * (1 + 1e-9)) - L.oracle_xlogx(s * (1 - 1e-9))) < 1e-15


def test_interpolation_functions(L):
    """testInterpolationFunctions.cc:23-63"""
    tol = 1e-8
    assert abs(L.oracle_interp_func(0.5, b"p") - 0.5) < tol
    assert abs(L.oracle_interp_func(-0.5, b"p")) < tol
    assert abs(L.oracle_interp_func(1.5, b"p") - 1.0) < tol
    for phi in (0.0, 0.2, 0.5, 0.9, 1.0):
        assert abs(L.oracle_interp_ratio_func(phi, b"p", b"p") - 1.0) < tol
        # ratio pbg/lin = phi^2 (10 - 15 phi + 6 phi^2)   (reference test, phi = 0.05)
        r = L.oracle_interp_ratio_func(phi, b"p", b"l")
        assert abs(r - phi * phi * (10.0 - 15.0 * phi + 6 * phi * phi)) < tol
        assert abs(r * phi - L.oracle_interp_func(phi, b"p")) < tol
        # compl ratio = (1 - p(phi)) / (1 - phi)
        cr = L.oracle_compl_interp_ratio_func(phi, b"p", b"l")
        assert abs(cr * (1 - phi) - (1 - L.oracle_interp_func(phi, b"p"))) < tol
    r = L.oracle_interp_ratio_func(0.05, b"p", b"l")
    assert abs(r - 0.05 * 0.05 * (10.0 - 15.0 * 0.05 + 6 * 0.05 * 0.05)) < tol
    a = 1.0 - L.oracle_interp_func(0.05, b"p")
    b = 1.0 - L.oracle_interp_func(0.05, b"l")
    assert abs(L.oracle_compl_interp_ratio_func(0.05, b"p", b"l") - a / b) < tol
    pass
    # quirks that the CUDA path must reproduce
    assert L.oracle_deriv_interp_func(-3.0, b"l") == 1.0
    assert L.oracle_deriv_interp_func(7.0, b"l") == 1.0
    assert L.oracle_average_func(1e-17, 0.5, b"h") == 0.0
    assert abs(L.oracle_average_func(0.25, 0.75, b"a") - 0.5) < 1e-16
    # derivative consistency
    for t in (b"q", b"p", b"h", b"w", b"m", b"3"):
        phi, e = 0.37, 1e-7
        fd = (L.oracle_interp_func(phi + e, t) - L.oracle_interp_func(phi - e, t)) / (2 * e)
        assert abs(fd - L.oracle_deriv_interp_func(phi, t)) < 1e-6


def _side_arrays(ndim, lo, hi, ng, depth):
    arrs = []
    for a in range(ndim):
        shape = [hi[d] - lo[d] + 1 + 2 * ng + (1 if d == a else 0) for d in range(ndim)]
        arrs.append(np.zeros([depth] + shape[::-1]))
    return arrs


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("qlen", [1, 2, 4])
def test_gradq_linear_field(L, ndim, qlen):
    """testGradQ.cc: q linear in x,y,z -> diffs = slope, cell / side gradients = slope/dx"""
    ng = 1
    lo = [0] * ndim
    hi = [(d + 2) * 3 for d in range(ndim)]
    dx = np.array([0.1, 0.11, 0.12])
    alpha = np.array([[(axis + 1) + 0.1 * q for q in range(qlen)] for axis in range(ndim)])
    shape = [hi[d] - lo[d] + 1 + 2 * ng for d in range(ndim)]
    idx = np.meshgrid(*[np.arange(s) for s in shape[::-1]], indexing="ij")  # (z,y,x) order
    quat = np.zeros([qlen] + shape[::-1])
    for q in range(qlen):
        for axis in range(ndim):
            quat[q] += idx[ndim - 1 - axis] * alpha[axis, q]
    diffs = _side_arrays(ndim, lo, hi, ng, qlen)
    ilo = (C.c_int * 3)(*lo, *([0] * (3 - ndim)))
    ihi = (C.c_int * 3)(*hi, *([0] * (3 - ndim)))
    pd = (C.c_void_p * 3)(*[a.ctypes.data for a in diffs], *([None] * (3 - ndim)))
    L.oracle_k_quatdiffs(ndim, ilo, ihi, qlen, quat.ctypes.data_as(C.c_void_p), ng, pd, ng)
    for axis in range(ndim):
        d = diffs[axis]
        # interior side box: strip the ghosts
        sl = tuple([slice(None)] + [slice(ng, -ng)] * ndim)
        for q in range(qlen):
            assert np.abs(d[sl][q] - alpha[axis, q]).max() < 1e-6
    # cell gradients
    gshape = [hi[d] - lo[d] + 1 for d in range(ndim)]
    grads = [np.zeros([qlen] + gshape[::-1]) for _ in range(ndim)]
    pg = (C.c_void_p * 3)(*[a.ctypes.data for a in grads], *([None] * (3 - ndim)))
    pdx = (C.c_double * 3)(*dx)
    L.oracle_k_quatgrad_cell(ndim, ilo, ihi, qlen, pdx, pd, ng, pg, 0)
    for axis in range(ndim):
        for q in range(qlen):
            assert np.abs(grads[axis][q] - alpha[axis, q] / dx[axis]).max() < 1e-6
    # side gradients: component dir*qlen + q
    gs = _side_arrays(ndim, lo, hi, 0, ndim * qlen)
    pgs = (C.c_void_p * 3)(*[a.ctypes.data for a in gs], *([None] * (3 - ndim)))
    L.oracle_k_quatgrad_side(ndim, ilo, ihi, qlen, pdx, pd, ng, pgs, 0)
    for axis in range(ndim):
        for dr in range(ndim):
            for q in range(qlen):
                assert np.abs(gs[axis][dr * qlen + q] - alpha[dr, q] / dx[dr]).max() < 1e-6


@pytest.mark.parametrize("ndim", [2, 3])
def test_add_flux_linear_field(L, ndim):
    """testFlux.cc pattern: linear field, D = 111 -> flux = D * slope / dx (tol 1e-6)"""
    lo = [0] * ndim
    hi = [6, 9, 12][:ndim]
    dx = np.array([0.2, 0.15, 0.1])
    slope = np.array([1.0, 2.0, 3.0])
    ng = 1
    shape = [hi[d] - lo[d] + 1 + 2 * ng for d in range(ndim)]
    idx = np.meshgrid(*[np.arange(s) for s in shape[::-1]], indexing="ij")
    conc = np.zeros([1] + shape[::-1])
    for axis in range(ndim):
        conc[0] += idx[ndim - 1 - axis] * slope[axis]
    D = _side_arrays(ndim, lo, hi, 0, 1)
    for a in D:
        a[...] = 111.0
    F = _side_arrays(ndim, lo, hi, 0, 1)
    ilo = (C.c_int * 3)(*lo, *([0] * (3 - ndim)))
    ihi = (C.c_int * 3)(*hi, *([0] * (3 - ndim)))
    pD = (C.c_void_p * 3)(*[a.ctypes.data for a in D], *([None] * (3 - ndim)))
    pF = (C.c_void_p * 3)(*[a.ctypes.data for a in F], *([None] * (3 - ndim)))
    pdx = (C.c_double * 3)(*dx)
    L.oracle_k_add_flux(ndim, ilo, ihi, pdx, conc.ctypes.data_as(C.c_void_p), ng, 1, pD, 0, pF, 0)
    for axis in range(ndim):
        assert np.abs(F[axis] - 111.0 * slope[axis] / dx[axis]).max() < 1e-6


def test_symmetry_rotation_table(L):
    """setqr (quat.f:165-286): 48 unit quaternions, conjugate table really conjugates"""
    tab = np.zeros((48, 4))
    L.oracle_qr_table4(tab.ctypes.data)
    assert np.abs(np.linalg.norm(tab, axis=1) - 1.0).max() < 1e-15
    q = np.array([0.3, -0.5, 0.1, 0.8])
    q /= np.linalg.norm(q)
    for iq in range(1, 49):
        qp = np.zeros(4)
        back = np.zeros(4)
        L.oracle_quatsymmrotate(q.ctypes.data, iq, qp.ctypes.data, 4)
        L.oracle_quatsymmrotate(qp.ctypes.data, -iq, back.ctypes.data, 4)
        assert np.abs(back - q).max() < 1e-14
    for iq in range(1, 5):
        q2 = np.array([np.cos(0.4), np.sin(0.4)])
        qp = np.zeros(2)
        back = np.zeros(2)
        L.oracle_quatsymmrotate(q2.ctypes.data, iq, qp.ctypes.data, 2)
        L.oracle_quatsymmrotate(qp.ctypes.data, -iq, back.ctypes.data, 2)
        assert np.abs(back - q2).max() < 1e-15


def test_grad_normi_floor(L):
    f, mx = 1e-2, 1e2
    assert L.oracle_eval_grad_normi(4.0, b"m", f * f, mx) == 0.5
    assert L.oracle_eval_grad_normi(1e-8, b"m", f * f, mx) == mx
    g2 = 0.3
    assert abs(L.oracle_eval_grad_normi(g2, b"s", f * f, mx) - 1 / np.sqrt(g2 + f * f)) < 1e-15
    assert abs(L.oracle_eval_grad_normi(g2, b"t", f * f, mx) - np.tanh(np.sqrt(g2) / f) / np.sqrt(g2)) < 1e-14
    # Taylor branch continuous with the tanh branch at x = 0.01
    a = L.oracle_eval_grad_normi(0.0099999 * f * f, b"t", f * f, mx)
    b = L.oracle_eval_grad_normi(0.0100001 * f * f, b"t", f * f, mx)
    assert abs(a - b) / a < 1e-5
