"""Run an AMPE input deck on the device:   python -m ampe_b200.run_deck 2d.input [--precond-cycles N]

The run loop around the path (PFModel::Run, source/PFModel.cc:430-560; QuatModel::postAdvanceDiagnostics /
preRunDiagnostics / postRunDiagnostics, QuatModel.cc:2369-2373, 2431-2433, 2537-2541; EventInterval.cc): the deck is
read by ampe_b200.input_deck, the initial conditions by FieldsInitializer (NetCDF classic or NetCDF-4 container, the
uniform `init_t` / `init_q` / `init_c` of InitialConditions{} on top), the state is integrated by the implicit integrator
on the device under the deck's tolerances, and at the events of ScalarDiagnostics{} / GrainDiagnostics{} the lines the
reference prints are printed in the reference's format --

   cycle # 212 : t = 0.3002 : dt = 0.0021
     Volume fraction of solid phase = 0.3189
     Integral concentration 0= 0.8099
     Volume of grain 3 = 1.9994

-- so the analysis scripts of the reference's regression tests (tests/*/test2d.py: words[6] of the "cycle" and
"fraction" lines, words[3] of the "Integral" lines) read this program's output as they read AMPE's.

Events.  AMPE advances one CVODE step per cycle and tests its intervals after each; here the integrator runs from one
event to the next (`stop_at_tend` off: it returns after the first step AT OR BEYOND the event time, so an output lands
where AMPE's lands), which restarts the BDF history at every event.  Intervals counted in steps are run as blocks of
that many steps.  The "cycle" line is printed at the events, not at every step.

The loop itself is host logic over a small backend interface (integrate / diagnostics / grains); the only backend of
the product is the device (`DeviceBackend`); tests/test_run_deck.py drives the same loop with the CPU restatement.
"""
import argparse
import os
import sys

import numpy as np

from . import input_deck

IMPLICIT_ETOOMUCHWORK = -22


class EventInterval:
    """EventInterval.cc: an interval in steps or in time, optionally firing at the first and the last step"""

    def __init__(self, db, name, default_value=0.0, default_type="step", include_first=False, include_last=True):
        blk = db.get(name) if isinstance(db.get(name), dict) else {}
        kind = str(blk.get("interval_type", default_type))
        if kind not in ("step", "time", "dt", "cycle"):
            raise input_deck.DeckError('Error in EventInterval "%s": invalid interval_type' % name)
        self.by_step = kind in ("step", "cycle")
        self.include_first = bool(blk.get("include_first_step", include_first))
        self.include_last = bool(blk.get("include_last_step", include_last))
        self.steps = int(blk.get("interval", int(default_value + 0.5))) if self.by_step else 0
        self.dt = 0.0 if self.by_step else float(blk.get("interval", default_value))
        self.previous_time, self.time_at_last_event = -1.0, -1.0

    def active(self):
        return self.steps > 0 if self.by_step else self.dt > 0.0

    def _fresh(self, t):
        return self.active() and self.time_at_last_event != t

    def include_initial(self, t):
        fire = self._fresh(t) and self.include_first
        if fire:
            self.time_at_last_event = t
        self.previous_time = t
        return fire

    def include_final(self, t):
        fire = self._fresh(t) and self.include_last
        if fire:
            self.time_at_last_event = t
        self.previous_time = t
        return fire

    def has_interval_passed(self, step, t):
        fire = False
        if self._fresh(t):
            if self.include_first and self.previous_time < 0.0:
                fire = True
            elif self.by_step:
                fire = step % self.steps == 0
            elif self.previous_time >= 0.0:
                fire = int(t / self.dt) > int(self.previous_time / self.dt)
        self.previous_time = t
        if fire:
            self.time_at_last_event = t
        return fire

    def next_stop(self, step, t):
        """(time, steps) of the next possible event: where the integrator is asked to pause"""
        if not self.active():
            return None, None
        if self.by_step:
            return None, self.steps - step % self.steps
        return (int(t / self.dt) + 1) * self.dt, None


def initial_state(db, cfg, deck_dir, read_initial_conditions):
    """InitialConditions{} (PFModel.cc:351-399): the fields of the file, overridden by the uniform init_* values"""
    run = input_deck.run_parameters(db)
    nz = cfg.n[2] if cfg.ndim == 3 else 1
    shape = (nz, cfg.n[1], cfg.n[0])
    fields = []
    if cfg.with_phase:
        fields.append("phase")
    if cfg.qlen > 0 and run["init_q"] is None:
        fields.append("quat")
    if cfg.with_concentration and run["init_c"] is None:
        fields.append("conc")
    if cfg.with_unsteady_temperature and run["init_t"] is None:
        fields.append("temperature")
    y = {"phase": None, "quat": None, "conc": None, "temperature": None}
    name = run["initial_conditions_file"]
    if name is not None:
        path = name if os.path.isabs(name) or os.path.exists(name) else os.path.join(deck_dir, name)
        got = read_initial_conditions(path, cfg, slice_index=run["slice_index"], fields=tuple(fields))
        y.update({k: (None if v is None else np.array(v, dtype=np.float64, copy=True)) for k, v in got.items() if k in fields})
    elif fields:
        raise input_deck.DeckError("InitialConditions{filename} is required for %s" % ", ".join(fields))
    if cfg.qlen > 0 and run["init_q"] is not None:
        q = np.asarray(run["init_q"], dtype=np.float64)
        if q.size != cfg.qlen:
            raise input_deck.DeckError("InitialConditions{init_q} needs %d values" % cfg.qlen)
        y["quat"] = np.ascontiguousarray(np.broadcast_to(q[:, None, None, None], (cfg.qlen,) + shape))
    if cfg.with_concentration and run["init_c"] is not None:
        y["conc"] = np.full(shape, float(run["init_c"] if not isinstance(run["init_c"], list) else run["init_c"][0]))
    if cfg.with_unsteady_temperature and run["init_t"] is not None:
        y["temperature"] = np.full(shape, float(run["init_t"]))
    return y


def print_scalar_diagnostics(cfg, d, out):
    """the lines of QuatModel::printScalarDiagnostics (QuatModel.cc:2543-2690) the fused path has numbers for"""
    if cfg.with_unsteady_temperature:
        out.write("Thermal energy [pJ]= %.8g\n" % d["thermal_energy"])
        out.write("  Min. Temperature = %.8g\n  Max. Temperature = %.8g\n  Average Temperature = %.8g\n"
                  % (d["min_temperature"], d["max_temperature"], d["average_temperature"]))
    if cfg.with_phase:
        out.write("  Volume fraction of solid phase = %.8g\n" % d["solid_fraction"])
    if cfg.with_concentration:
        out.write("  Integral concentration 0= %.8g\n  Max. concentration 0= %.8g\n" % (d["integral_concentration"], d["max_concentration"]))
        if cfg.with_phase:
            out.write("  Cex (HBSM) for component 0 = %.8g\n" % d["cex"])


def run(db, cfg, y, backend, h0=None, out=None):
    """PFModel::Run: returns (cycles, time, history of (cycle, time, diagnostics))"""
    out = sys.stdout if out is None else out
    par = input_deck.run_parameters(db)
    if par["end_time"] is None:
        raise input_deck.DeckError("key 'end_time' is required")
    end_time, max_cycles = par["end_time"], par["max_timesteps"]
    scalar = EventInterval(db, "ScalarDiagnostics", 0.0, "step")
    grain = EventInterval(db, "GrainDiagnostics", 0.0, "step")
    gd = db.get("GrainDiagnostics") if isinstance(db.get("GrainDiagnostics"), dict) else {}
    threshold = float(gd.get("phase_threshold", 0.85))
    history = []

    def diagnostics(cycle, t, dt, scalars, grains):
        out.write("cycle # %d : t = %.10g%s\n" % (cycle, t, "" if dt is None else " : dt = %.6g" % dt))
        rec = {}
        if scalars:
            rec = backend.scalar_diagnostics(y)
            print_scalar_diagnostics(cfg, rec, out)
        if grains:
            out.write("findAndNumberGrains\n")
            vols = backend.grain_volumes(y, threshold)
            rec = dict(rec, grain_volumes=vols)
            for g in sorted(vols):
                out.write("Volume of grain %d = %.8g\n" % (g, vols[g]))
        history.append((cycle, t, rec))
        out.flush()

    t, cycle = 0.0, 0
    h = h0 if h0 is not None else 1.0e-6 * min(end_time, 1.0)   # the first step; the controller takes over at once
    s0, g0 = scalar.include_initial(t), grain.include_initial(t)
    if s0 or g0:
        diagnostics(cycle, t, None, s0, g0)
    while t < end_time and cycle < max_cycles:
        stop_t, stop_n = end_time, max_cycles - cycle
        for ev in (scalar, grain):
            et, en = ev.next_stop(cycle, t)
            if et is not None:
                stop_t = min(stop_t, et)
            if en is not None:
                stop_n = min(stop_n, en)
        rc, st = backend.integrate(y, stop_t, h, t, par["rtol"], par["atol"], stop_n)
        if rc not in (0, IMPLICIT_ETOOMUCHWORK) or int(st["steps"]) == 0:
            raise RuntimeError("the integrator failed (rc %d) at t = %g: %r" % (rc, st.get("t_reached", t), st))
        t, h, cycle = st["t_reached"], st["last_step"], cycle + int(st["steps"])
        s1, g1 = scalar.has_interval_passed(cycle, t), grain.has_interval_passed(cycle, t)
        if s1 or g1:
            diagnostics(cycle, t, h, s1, g1)
    s2, g2 = scalar.include_final(t), grain.include_final(t)
    if s2 or g2:
        diagnostics(cycle, t, h, s2, g2)
    return cycle, t, history


def current_temperature(cfg, t):
    """ScalarTemperatureStrategy::getCurrentTemperature (ScalarTemperatureStrategy.cc:57-74): the ramp, held at its target"""
    T = cfg.T_uniform + cfg.dtemperaturedt * t
    if cfg.dtemperaturedt < 0.0 and T < cfg.target_temperature:
        return cfg.target_temperature
    if cfg.target_temperature > 0.0 and cfg.dtemperaturedt > 0.0 and T > cfg.target_temperature:
        return cfg.target_temperature
    return T


def write_ending_file(db, cfg, y_np, t, directory="."):
    """InitialConditions{WriteEndingFile{filename}} (PFModel.cc:146-160, FieldsWriter.cc:60-330): the fields at the end of the
    run as single-precision variables phase / quat1.. / concentration / temperature dimensioned (z, y, x) plus the dimension
    qlen -- the file a follow-up deck names as its InitialConditions{filename} (examples/AuNi_2D: 9grains_AuNi_initial.input,
    then 9grains_AuNi.input).  Written as NetCDF classic, the HAVE_NETCDF3 branch of the reference.  Returns the path or None."""
    ic = db.get("InitialConditions") if isinstance(db.get("InitialConditions"), dict) else {}
    end = ic.get("WriteEndingFile")
    if not isinstance(end, dict):
        return None
    if "filename" not in end:
        raise input_deck.DeckError("key 'filename' is required in WriteEndingFile")
    from . import netcdf_classic
    nz = cfg.n[2] if cfg.ndim == 3 else 1
    state = dict(y_np)
    if state.get("temperature") is None:
        state["temperature"] = np.full((nz, cfg.n[1], cfg.n[0]), current_temperature(cfg, t))
    name = str(end["filename"])
    path = name if os.path.isabs(name) else os.path.join(directory, name)
    netcdf_classic.write_state(path, state, qlen=cfg.qlen, dtype=np.float32)
    return path


class DeviceBackend:
    """the product's backend: state and integrator on the GPU (host/QuatIntegrator.h through the C ABI)"""

    def __init__(self, cfg, precond_cycles=0, scale_newton_tolerance=False, precondition_left=False, strict_linear=False,
                 hold_step_after_failure=False):
        from . import host_rhs, rhs
        self.cfg, self._rhs, self.scale_newton_tolerance, self.strict_linear = cfg, rhs, scale_newton_tolerance, strict_linear
        self.hold_step_after_failure = hold_step_after_failure
        self.integrator = host_rhs.HostQuatIntegrator(cfg, True)
        if precond_cycles:
            self.integrator.setupPreconditioners(precond_cycles, precondition_left=precondition_left)
        self.diag = rhs.QuatIntegratorRHS(cfg)

    def upload(self, y_np):
        import torch
        y = self._rhs.SolutionVector({k: (None if v is None else torch.as_tensor(np.ascontiguousarray(v)).cuda())
                                      for k, v in y_np.items()})
        if self.cfg.conc_rhs_form in (2, 3):   # the Newton solves start from the composition itself (QuatModel.cc:1290-1330)
            c0 = y["conc"].reshape(-1).clone()
            self.integrator.resetRefPhaseConcentrations(c0, c0.clone())
        return y

    def integrate(self, y, tend, h, t0, rtol, atol, max_steps):
        return self.integrator.integrateAdaptive(y, tend, h, t0=t0, rtol=rtol, atol=atol, max_steps=max_steps, stop_at_tend=False,
                                                 scale_newton_tolerance=self.scale_newton_tolerance, strict_linear=self.strict_linear,
                                                 hold_step_after_failure=self.hold_step_after_failure)

    def scalar_diagnostics(self, y):
        return self.diag.printScalarDiagnostics(y)

    def grain_volumes(self, y, threshold):
        return self.diag.computeGrainDiagnostics(y, threshold)

    def download(self, y):
        return {k: (None if v is None else v.cpu().numpy()) for k, v in y.items()}

    def close(self):
        self.integrator.close()
        self.diag.close()


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("deck")
    ap.add_argument("--precond-cycles", type=int, default=2, help="V-cycles of the block preconditioners (0: none)")
    ap.add_argument("--first-step", type=float, default=None)
    ap.add_argument("--cvode-newton-tolerance", action="store_true",
                    help="bound the Newton error by nlscoef of the allowed local error as CVODE does (ImplicitOptions::scale_newton_tolerance)")
    ap.add_argument("--precondition-left", action="store_true",
                    help="left preconditioning as AMPE configures CVODE (PREC_LEFT, QuatIntegrator.cc:1583): GMRES tests the "
                         "preconditioned residual, which is what lets very stiff decks (tests/ConservedVolume) take large steps")
    ap.add_argument("--cvode-linear-rule", action="store_true",
                    help="an unconverged linear solve is accepted on the first Newton iteration only (ImplicitOptions::strict_linear_convergence)")
    ap.add_argument("--cvode-hold-step", action="store_true",
                    help="no step growth on the step that follows a failed attempt, CVODE's etamax = 1 (ImplicitOptions::hold_step_after_failure)")
    a = ap.parse_args(argv)
    from . import host_rhs
    db = input_deck.load(a.deck)
    cfg = input_deck.rhs_config(db, deck_dir=os.path.dirname(os.path.abspath(a.deck)))
    y_np = initial_state(db, cfg, os.path.dirname(os.path.abspath(a.deck)),
                         lambda *args, **kw: {k: (None if v is None else v.numpy()) for k, v in
                                              host_rhs.read_initial_conditions(*args, **kw).items()})
    backend = DeviceBackend(cfg, a.precond_cycles, a.cvode_newton_tolerance, a.precondition_left, a.cvode_linear_rule, a.cvode_hold_step)
    try:
        y = backend.upload(y_np)
        cycles, t, _ = run(db, cfg, y, backend, h0=a.first_step)
        written = write_ending_file(db, cfg, backend.download(y), t)
        if written:
            print("Open/replace file %s" % written)
    finally:
        backend.close()
    print("Run complete: %d steps, end time %.10g" % (cycles, t))
    return 0


if __name__ == "__main__":
    sys.exit(main())
