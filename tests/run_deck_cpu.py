"""Test infrastructure: `python tests/run_deck_cpu.py deck.input` = ampe_b200.run_deck's loop with the CPU restatement as the
backend, so the reference's test scripts can be pointed at an executable in a container without a GPU (the product's own
program, `python -m ampe_b200.run_deck`, has the device as its only backend).  AMPE_B200_CPU_DECK_OPTS, a comma-separated list of
left, strict, scale, hold, cycles=N, selects the integrator options run_deck's command line offers for the device."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ampe_b200 import host_rhs, input_deck, run_deck  # noqa: E402
from test_run_deck import OracleBackend, _read  # noqa: E402

if __name__ == "__main__":
    deck = sys.argv[1]
    db = input_deck.load(deck)
    cfg = input_deck.rhs_config(db)
    y = run_deck.initial_state(db, cfg, os.path.dirname(os.path.abspath(deck)), _read)
    opts = [o for o in os.environ.get("AMPE_B200_CPU_DECK_OPTS", "").split(",") if o]
    ncyc = ([int(o[len("cycles="):]) for o in opts if o.startswith("cycles=")] or [2])[0]
    backend = OracleBackend(cfg, y, precond_cycles=0 if opts else ncyc)
    if opts:
        backend.o.set_preconditioner(ncyc, left="left" in opts)
        backend.integrate = lambda y, tend, h, t0, rtol, atol, max_steps: backend.o.integrate_adaptive(
            y, tend, h, t0=t0, rtol=rtol, atol=atol, max_steps=max_steps, stop_at_tend=False, strict_linear="strict" in opts,
            scale_newton_tolerance="scale" in opts, hold_step_after_failure="hold" in opts)
    try:
        cycles, t, _ = run_deck.run(db, cfg, y, backend)
        written = run_deck.write_ending_file(db, cfg, y, t)
        if written:
            print("Open/replace file %s" % written)
    finally:
        backend.close()
    print("Run complete: %d steps, end time %.10g" % (cycles, t))
