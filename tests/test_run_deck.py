"""The run loop around the path (ampe_b200/run_deck.py, mirror of PFModel::Run and the diagnostics events of QuatModel)
on the CPU: the loop is host logic over a backend; here the backend is the CPU restatement (test infrastructure), on the
device it is the product's integrator (`python -m ampe_b200.run_deck deck.input`).  The decisive case takes a regression
deck of the reference as the FILE the reference ships, the initial condition its generator writes packed as a NetCDF-4
container under the name the deck asks for, and holds the program's standard output against the acceptance logic of the
reference's own tests/OneGrainQuadratic/test2d.py (which parses AMPE's output: words[6] of the "cycle" and "fraction" lines)."""
import io
import os

import numpy as np
import pytest

import hdf5_writer
from ampe_b200 import host_rhs, input_deck, run_deck

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    def __init__(self, cfg, y, precond_cycles=0):
        from oracle import pyoracle
        self.o = pyoracle.Oracle(cfg, perf=True)
        self.o.L.oracle_set_num_threads(len(os.sched_getaffinity(0)))
        if cfg.conc_rhs_form in (2, 3):
            self.o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
        if precond_cycles:
            self.o.set_preconditioner(precond_cycles)
        self.calls = []

    def integrate(self, y, tend, h, t0, rtol, atol, max_steps):
        self.calls.append((t0, tend, max_steps))
        return self.o.integrate_adaptive(y, tend, h, t0=t0, rtol=rtol, atol=atol, max_steps=max_steps, stop_at_tend=False)

    def scalar_diagnostics(self, y):
        return self.o.scalar_diagnostics(y)

    def grain_volumes(self, y, threshold):
        return self.o.grain_volumes(y, threshold)

    def close(self):
        self.o.close()


def _read(path, cfg, **kw):
    return {k: (None if v is None else v.numpy()) for k, v in host_rhs.read_initial_conditions(path, cfg, **kw).items()}


def test_event_interval_follows_the_reference():
    """EventInterval.cc: time intervals fire when int(t / dt) grows, step intervals on multiples; an event is not repeated at the
    same time; include_first_step / include_last_step"""
    db = input_deck.parse('A { interval = 0.05 interval_type = "time" } B { interval = 20 interval_type = "step" include_first_step = TRUE }'
                          ' C { interval = 0.1 interval_type = "time" include_last_step = FALSE }')
    a, b, c = (run_deck.EventInterval(db, k) for k in "ABC")
    none = run_deck.EventInterval(db, "Missing", 0.0, "step")
    assert a.active() and b.active() and not none.active()
    assert not a.include_initial(0.0) and b.include_initial(0.0) and not none.include_initial(0.0)
    assert not a.has_interval_passed(1, 0.03) and a.has_interval_passed(2, 0.051) and not a.has_interval_passed(3, 0.09)
    assert a.has_interval_passed(4, 0.21)                       # several intervals in one step: one event
    assert not a.has_interval_passed(5, 0.21)                   # not twice at the same time
    assert not b.has_interval_passed(19, 0.1) and b.has_interval_passed(20, 0.2) and b.has_interval_passed(40, 0.3)
    assert a.next_stop(5, 0.21) == (0.25, None) and b.next_stop(45, 0.3) == (None, 15)
    assert a.include_final(0.26) and not a.include_final(0.26)
    assert not c.include_final(0.26)
    with pytest.raises(input_deck.DeckError, match="invalid interval_type"):
        run_deck.EventInterval(input_deck.parse('A { interval_type = "often" }'), "A")


DECK = '''
// a small solidification problem written for this test: one quadratic-energy grain in an undercooled melt
end_time = 2.e-3
max_timesteps = 200
ModelParameters {
   Temperature { type = "scalar"  temperature = 873.  dtemperaturedt = -20.  target_temperature = 573. }
   Interface { sigma = 0.1  delta = 0.045 }
   phi_mobility = 200.
   phi_interp_func_type = "harmonic"
   avg_func_type = "arithmetic"
   ConcentrationModel {
      model = "quadratic"   rhs_form = "ebs"   diffusion_type = "temperature_dependent"
      molar_volume = 1.5e-5
      D_solid = 1.3e8  Q0_solid = 156377.   D_liquid = 5.6e4  Q0_liquid = 55329.
      Quadratic { T_ref = 1000.  A_liquid = 1.e4  A_solid = 1.e4  Ceq_liquid = 0.05  Ceq_solid = 0.1  m_liquid = 0.  m_solid = 0. }
   }
}
Integrator { atol = 1.e-4 }
ScalarDiagnostics { interval = 5.e-4  interval_type = "time"  include_first_step = TRUE }
GrainDiagnostics { interval = 1.e-3  interval_type = "time"  phase_threshold = 0.5 }
InitialConditions { filename = "disc.nc" }
Geometry { coarsest_level_resolution = 32, 32   x_lo = 0., 0.   x_up = 1.6, 1.6 }
'''


def test_run_loop_on_a_deck_written_here(tmp_path):
    """events of two intervals, the first and the last step, the reference's line formats, a NetCDF-4 initial condition"""
    db = input_deck.parse(DECK)
    cfg = input_deck.rhs_config(db)
    j, i = np.meshgrid(np.arange(32) + 0.5, np.arange(32) + 0.5, indexing="ij")
    r = np.sqrt((i - 16.0) ** 2 + (j - 16.0) ** 2)
    phase = 0.5 * (1.0 - np.tanh((r - 8.0) / 1.5))
    conc = 0.06 + 0.04 * phase
    hdf5_writer.write_hdf5(str(tmp_path / "disc.nc"), {"phase": phase[None].astype(np.float32), "concentration0": conc[None].astype(np.float32)},
                           dimensions={"x": 32, "y": 32, "z": 1})
    y = run_deck.initial_state(db, cfg, str(tmp_path), _read)
    assert np.array_equal(y["phase"][0], phase.astype(np.float32).astype(np.float64)) and y["quat"] is None
    backend = OracleBackend(cfg, y, precond_cycles=2)
    out = io.StringIO()
    try:
        cycles, t, hist = run_deck.run(db, cfg, y, backend, out=out)
    finally:
        backend.close()
    text = out.getvalue()
    lines = text.splitlines()
    assert t >= 2.0e-3 and cycles <= 200
    times = [h[1] for h in hist]
    assert times[0] == 0.0 and len(hist) == 5                                  # t = 0 and the four multiples of 5e-4
    for k, tk in enumerate(times[1:], 1):
        assert 5.0e-4 * k <= tk < 5.0e-4 * (k + 1)                             # each output on the first step at or past its time
    assert [("grain_volumes" in h[2]) for h in hist] == [False, False, True, False, True]
    assert all(len(h[2]["grain_volumes"]) == 1 for h in hist if "grain_volumes" in h[2])   # one grain
    assert lines[0] == "cycle # 0 : t = 0" and lines[1].startswith("  Volume fraction of solid phase = ")
    cyc = [ln for ln in lines if "cycle" in ln]
    assert len(cyc) == 5 and all(abs(float(ln.split()[6]) - tk) <= 1.0e-9 * tk for ln, tk in zip(cyc, times))   # words[6] is the time
    frac = [float(ln.split()[6]) for ln in lines if "fraction" in ln]
    assert len(frac) == 5 and all(f2 >= f1 for f1, f2 in zip(frac, frac[1:]))    # the undercooled grain grows
    integral = [float(ln.split()[3]) for ln in lines if "Integral" in ln]
    assert len(integral) == 5 and max(integral) - min(integral) <= 1.0e-4 * integral[0]   # words[3]; composition is conserved
    assert sum("Volume of grain" in ln for ln in lines) == 2
    # the integrator was asked to pause at the event times, never beyond the end
    assert [round(c[1] / 5.0e-4) for c in backend.calls] == [1, 2, 3, 4]


@pytest.mark.skipif(not os.path.exists(REF + "/tests/OneGrainQuadratic/2d.input"), reason="runs the reference's own deck file (build container only)")
def test_reference_deck_file_to_the_reference_acceptance(tmp_path):
    """tests/OneGrainQuadratic: 2d.input as shipped; nuclei.nc = the arrays utils/make_nuclei.py writes for the test's command
    line, in a NetCDF-4 container; the output parsed as tests/OneGrainQuadratic/test2d.py:26-52 parses AMPE's"""
    db = input_deck.load(REF + "/tests/OneGrainQuadratic/2d.input")
    cfg = input_deck.rhs_config(db)
    ic = np.load(os.path.join(ROOT, "tests", "golden", "ic_one_grain_quadratic2d.npz"))
    name = input_deck.run_parameters(db)["initial_conditions_file"]
    hdf5_writer.write_hdf5(str(tmp_path / name), {k: ic[k] for k in ic.files}, dimensions={"x": 64, "y": 64, "z": 1})
    y = run_deck.initial_state(db, cfg, str(tmp_path), _read)
    backend = OracleBackend(cfg, y, precond_cycles=2)
    out = io.StringIO()
    try:
        cycles, t, _ = run_deck.run(db, cfg, y, backend, out=out)
    finally:
        backend.close()
    assert cycles <= 220                       # max_timesteps of the deck
    end_reached, checked = False, 0
    for line in out.getvalue().encode().split(b"\n"):
        if line.count(b"cycle"):
            if eval(line.split()[6]) > 0.25:
                end_reached = True
        if line.count(b"fraction") and end_reached:
            assert abs(eval(line.split()[6]) - 0.21) <= 1.0e-2, line
            checked += 1
    assert end_reached and checked >= 1


def _disc_problem(tmp_path):
    j, i = np.meshgrid(np.arange(32) + 0.5, np.arange(32) + 0.5, indexing="ij")
    r = np.sqrt((i - 16.0) ** 2 + (j - 16.0) ** 2)
    phase = 0.5 * (1.0 - np.tanh((r - 8.0) / 1.5))
    hdf5_writer.write_hdf5(str(tmp_path / "disc.nc"), {"phase": phase[None].astype(np.float32),
                                                       "concentration0": (0.06 + 0.04 * phase)[None].astype(np.float32)},
                           dimensions={"x": 32, "y": 32, "z": 1})
    deck = tmp_path / "disc.input"
    deck.write_text(DECK)
    return str(deck)


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_run_deck_program_on_the_device(tmp_path, capsys):
    """`python -m ampe_b200.run_deck disc.input`: deck file + NetCDF-4 initial condition -> device integration -> AMPE's output
    lines; the same deck through the same loop on the CPU restatement lands on the same solid fraction"""
    deck = _disc_problem(tmp_path)
    assert run_deck.main([deck]) == 0
    lines = capsys.readouterr().out.splitlines()
    cyc = [ln.split() for ln in lines if ln.startswith("cycle #")]
    frac = [float(ln.split()[6]) for ln in lines if "fraction" in ln]
    integral = [float(ln.split()[3]) for ln in lines if "Integral" in ln]
    assert len(cyc) == 5 and float(cyc[-1][6]) >= 2.0e-3 and int(cyc[-1][2]) <= 200
    assert len(frac) == 5 and max(integral) - min(integral) <= 1.0e-4 * integral[0]
    assert sum("Volume of grain" in ln for ln in lines) == 2
    db = input_deck.parse(DECK)
    cfg = input_deck.rhs_config(db)
    y = run_deck.initial_state(db, cfg, str(tmp_path), _read)
    backend = OracleBackend(cfg, y, precond_cycles=2)
    try:
        cycles, t, hist = run_deck.run(db, cfg, y, backend, out=io.StringIO())
    finally:
        backend.close()
    assert abs(hist[-1][2]["solid_fraction"] - frac[-1]) <= 2.0e-4, (hist[-1][2]["solid_fraction"], frac[-1])
    assert abs(cycles - int(cyc[-1][2])) <= 3


def reference_tree(tmp_path, deck_dir):
    """tmp/tests/<deck_dir> as the working directory, tmp/utils -> the reference's generators: the layout its test scripts assume"""
    (tmp_path / "tests" / deck_dir).mkdir(parents=True)
    os.symlink(REF + "/utils", str(tmp_path / "utils"))
    return str(tmp_path / "tests" / deck_dir)


@pytest.mark.skipif(not os.path.exists(REF + "/tests/OneGrainQuadratic/test2d.py"), reason="runs the reference's own test scripts (build container only)")
@pytest.mark.timeout(600)
@pytest.mark.parametrize("deck,script,init_file", [("OneGrainQuadratic", "test2d.py", "nuclei.nc"),
                                                   ("CahnHilliard", "test2d.py", "64x64.nc"), ("CahnHilliard", "test3d.py", "32x32x32.nc")])
def test_reference_test_script_unmodified(tmp_path, deck, script, init_file):
    """tests/<deck>/test{2,3}d.py run as the reference's CTest runs them -- `test2d.py <mpiexec> <-n> <1> <exe> <input>` -- with
    nothing of them changed: each calls the reference's generator (utils/make_nuclei.py, tests/CahnHilliard/make_initial.py, which
    find the netCDF4 stand-in and write a NetCDF-4 container), starts the executable it is given on the deck, parses the output and
    exits 0 when its acceptance holds (solid fraction 0.21 +- 0.01 within 220 steps; the integral of the composition constant to 1e-5
    relative over a spinodal decomposition in 2D and in 3D).  The executable here is the deck program with the CPU restatement behind
    it; on the device it is `python -m ampe_b200.run_deck` (profiles/r02ak_reference_test_scripts.log, r02al_*)."""
    import subprocess
    import sys
    cwd = reference_tree(tmp_path, deck)
    for f in os.listdir(os.path.join(REF, "tests", deck)):     # generators that live next to the deck (make_initial.py)
        if f.startswith("make_"):
            os.symlink(os.path.join(REF, "tests", deck, f), os.path.join(cwd, f))
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tools", "netcdf4_shim") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    exe = "%s %s" % (sys.executable, os.path.join(ROOT, "tests", "run_deck_cpu.py"))
    dim = script[len("test"):-len(".py")]
    r = subprocess.run([sys.executable, os.path.join(REF, "tests", deck, script), "", "", "", exe, os.path.join(REF, "tests", deck, dim + ".input")],
                       cwd=cwd, env=env, capture_output=True, text=True, timeout=500)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "cycle" in r.stdout and not os.path.exists(os.path.join(cwd, init_file))   # the script removes its file at the end


def test_cvode_nonlinear_tolerance_option(tmp_path):
    """ImplicitOptions::scale_newton_tolerance (CVODE's tq[4] = nlscoef / tq[2]: the Newton error bounded by nlscoef of the allowed
    LOCAL ERROR instead of nlscoef in the WRMS norm): the same trajectory within the integration tolerance for less nonlinear work"""
    db = input_deck.parse(DECK)
    cfg = input_deck.rhs_config(db)
    _disc_problem(tmp_path)
    out = {}
    for flag in (False, True):
        y = run_deck.initial_state(db, cfg, str(tmp_path), _read)
        backend = OracleBackend(cfg, y, precond_cycles=2)
        try:
            rc, st = backend.o.integrate_adaptive(y, 2.0e-3, 2.0e-9, rtol=1.0e-6, atol=1.0e-4, max_steps=500, scale_newton_tolerance=flag)
            assert rc == 0, (rc, st)
            out[flag] = (st, backend.o.scalar_diagnostics(y)["solid_fraction"])
        finally:
            backend.close()
    assert abs(out[True][1] - out[False][1]) <= 2.0e-4
    assert out[True][0]["newton_iterations"] <= out[False][0]["newton_iterations"]
    assert out[True][0]["convergence_failures"] <= out[False][0]["convergence_failures"]
    assert out[True][0]["linear_iterations"] < out[False][0]["linear_iterations"]


def test_write_ending_file_feeds_the_next_deck(tmp_path):
    """InitialConditions{WriteEndingFile{filename}}: the first stage leaves the file the second stage starts from (the two-deck
    pattern of examples/AuNi_2D); single precision, the reference's variable names, the scalar temperature of the end time"""
    from scipy.io import netcdf_file
    _disc_problem(tmp_path)
    stage1 = DECK.replace("end_time = 2.e-3", "end_time = 5.e-4").replace(
        'InitialConditions { filename = "disc.nc" }', 'InitialConditions { filename = "disc.nc" WriteEndingFile { filename = "stage1.nc" } }')
    db = input_deck.parse(stage1)
    cfg = input_deck.rhs_config(db)
    y = run_deck.initial_state(db, cfg, str(tmp_path), _read)
    backend = OracleBackend(cfg, y, precond_cycles=2)
    try:
        cycles, t, _ = run_deck.run(db, cfg, y, backend, out=io.StringIO())
    finally:
        backend.close()
    path = run_deck.write_ending_file(db, cfg, y, t, str(tmp_path))
    assert path == str(tmp_path / "stage1.nc")
    f = netcdf_file(path, "r", mmap=False)
    assert set(f.variables) == {"phase", "concentration", "temperature"} and f.variables["phase"].data.dtype == np.dtype(">f4")
    assert np.array_equal(f.variables["phase"].data, y["phase"].astype(np.float32))
    assert np.all(f.variables["temperature"].data == np.float32(873.0 - 20.0 * t))
    f.close()
    stage2 = DECK.replace('filename = "disc.nc"', 'filename = "stage1.nc"')
    db2 = input_deck.parse(stage2)
    y2 = run_deck.initial_state(db2, cfg, str(tmp_path), _read)
    assert np.array_equal(y2["phase"], y["phase"].astype(np.float32).astype(np.float64))
    assert np.array_equal(y2["conc"], y["conc"].astype(np.float32).astype(np.float64))
    assert run_deck.write_ending_file(db2, cfg, y2, 0.0, str(tmp_path)) is None
    assert run_deck.current_temperature(cfg, 100.0) == 573.0       # held at the target


def test_left_preconditioning_as_the_reference_configures_cvode(tmp_path):
    """AMPE runs CVODE with PREC_LEFT (QuatIntegrator.cc:1583): GMRES then tests the PRECONDITIONED residual.  Same trajectory as the
    right-preconditioned default on a mild problem; on the very stiff tests/ConservedVolume deck it is the difference between
    ~10 000 and ~450 steps."""
    db = input_deck.parse(DECK)
    cfg = input_deck.rhs_config(db)
    _disc_problem(tmp_path)
    out = {}
    for left in (False, True):
        y = run_deck.initial_state(db, cfg, str(tmp_path), _read)
        backend = OracleBackend(cfg, y)
        backend.o.set_preconditioner(2, left=left)
        try:
            rc, st = backend.o.integrate_adaptive(y, 2.0e-3, 2.0e-9, rtol=1.0e-6, atol=1.0e-4, max_steps=500)
            assert rc == 0, (rc, st)
            out[left] = (st, backend.o.scalar_diagnostics(y))
        finally:
            backend.close()
    assert abs(out[True][1]["solid_fraction"] - out[False][1]["solid_fraction"]) <= 2.0e-4
    assert abs(out[True][1]["integral_concentration"] - out[False][1]["integral_concentration"]) <= 1.0e-6
    assert abs(out[True][0]["steps"] - out[False][0]["steps"]) <= 5


def test_step_held_after_a_failed_attempt(tmp_path):
    """ImplicitOptions::hold_step_after_failure, CVODE's etamax = 1 (cvHandleNFlag / cvDoErrorTest -> cvPrepareNextStep): the step that
    succeeds after a failed attempt keeps its size once.  Started with a first step far too large the integrator must fail, and
    with the rule on the accepted step after each failure is not grown: same answer, never more failures than without it."""
    db = input_deck.parse(DECK)
    cfg = input_deck.rhs_config(db)
    _disc_problem(tmp_path)
    out = {}
    for hold in (False, True):
        y = run_deck.initial_state(db, cfg, str(tmp_path), _read)
        backend = OracleBackend(cfg, y, precond_cycles=2)
        try:
            rc, st = backend.o.integrate_adaptive(y, 2.0e-3, 1.0e-3, rtol=1.0e-6, atol=1.0e-4, max_steps=500, hold_step_after_failure=hold)
            assert rc == 0, (rc, st)
            out[hold] = (st, backend.o.scalar_diagnostics(y))
        finally:
            backend.close()
    failures = {k: v[0]["error_test_failures"] + v[0]["convergence_failures"] for k, v in out.items()}
    assert failures[False] >= 1 and failures[True] >= 1, failures      # the oversized first step was refused
    assert failures[True] <= failures[False]
    assert abs(out[True][1]["solid_fraction"] - out[False][1]["solid_fraction"]) <= 2.0e-4
