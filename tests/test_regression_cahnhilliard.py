"""The reference's regression test tests/CahnHilliard/test2d.py, end to end on this stack (CPU backend here, the
device backend in tests/test_gpu_widening_zzz_precond.py): initial conditions written like
tests/CahnHilliard/make_initial.py (float32, variable `concentration0`), read by FieldsInitializer, integrated
with the variable-step implicit integrator under the deck's tolerance (Integrator{atol = 1.e-4}, rtol = 1e-2 atol)
to end_time = 300 with scalar diagnostics every 10 time units (ScalarDiagnostics{interval = 10.}); acceptance as
in test2d.py:37-39: two consecutive integrals of the composition differ by less than 1e-5 relative, and the end
time is reached.  Deck: tests/CahnHilliard/2d.input (64 x 64 cells, 100 um periodic box, mobility 5, ca 0.3,
cb 0.7, well_scale 5, kappa 2)."""
import numpy as np

from ampe_b200 import configs, host_rhs, netcdf_classic


def make_initial(nx, ny, Lx=100.0, Ly=100.0):
    """tests/CahnHilliard/make_initial.py:60-75 (nz = 1)"""
    c0, epsilon = 0.5, 0.01
    x = (np.arange(nx) + 0.5) * (Lx / nx)
    y = (np.arange(ny) + 0.5) * (Ly / ny)
    X, Y = np.meshgrid(x, y, indexing="xy")  # arrays indexed [j, i]
    t1 = np.cos(0.105 * X) * np.cos(0.11 * Y)
    t2 = np.cos(0.13 * X) * np.cos(0.087 * Y)
    t3 = np.cos(0.025 * X - 0.15 * Y) * np.cos(0.07 * X - 0.02 * Y)
    return (c0 + epsilon * (t1 + t2 * t2 + t3)).astype(np.float32)[None]


def deck():
    cfg = configs.pfhub1a(nx=64, ny=64)  # same model block as tests/CahnHilliard/2d.input
    for a in range(2):
        cfg.dx[a] = 100.0 / 64
    return cfg


def test_cahnhilliard_regression_deck(tmp_path):
    from oracle import pyoracle
    cfg = deck()
    path = str(tmp_path / "64x64.nc")
    netcdf_classic.write(path, {"concentration0": make_initial(64, 64)})
    y = {k: (None if v is None else v.numpy()) for k, v in host_rhs.read_initial_conditions(path, cfg).items()}
    assert y["conc"].shape == (1, 64, 64) and abs(float(y["conc"].mean()) - 0.5) < 0.01
    o = pyoracle.Oracle(cfg)
    atol = 1.0e-4
    old, t, h, steps = -1.0, 0.0, 1.0e-3, 0
    while t < 300.0:
        rc, st = o.integrate_adaptive(y, t + 10.0, h, t0=t, rtol=1e-2 * atol, atol=atol, max_steps=5000)
        assert rc == 0, st
        t, h, steps = st["t_reached"], st["last_step"], steps + int(st["steps"])
        conc = o.scalar_diagnostics(y)["integral_concentration"]
        if old < 0.0:
            old = conc
        assert abs(conc - old) <= 1.0e-5 * conc, (t, conc, old)  # test2d.py:37-39
        old = conc
    o.close()
    assert t >= 300.0  # "End time not reached" otherwise
    assert steps < 3000
    # spinodal decomposition has happened: the composition has left the neighbourhood of 0.5
    assert y["conc"].max() - y["conc"].min() > 0.3
