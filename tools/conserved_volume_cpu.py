"""The runs behind DESIGN.md 4's tests/ConservedVolume table: the reference's deck and its make_square.py initial condition through
run_deck's loop with the CPU restatement behind it (test infrastructure, like tests/run_deck_cpu.py; needs /root/reference, so
the build container only), once per (preconditioning side, V-cycles per block solve).  Prints the deck's two acceptance
quantities: the time reached within max_timesteps and the largest move of the solid fraction between consecutive outputs from
the fifth on (tests/ConservedVolume/test2d.py: <= 1e-4 and t > 0.01).

    python tools/conserved_volume_cpu.py [left:cycles ...]        e.g.  right:2 left:2 left:10 left:10:hold
(hold: ImplicitOptions::hold_step_after_failure, CVODE's etamax = 1 after a failed attempt)
"""
import io
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"

from ampe_b200 import input_deck, run_deck  # noqa: E402
from test_run_deck import OracleBackend, _read  # noqa: E402


def one(deck, cwd, left, cycles, hold=False):
    db = input_deck.load(deck)
    cfg = input_deck.rhs_config(db)
    y = run_deck.initial_state(db, cfg, cwd, _read)
    backend = OracleBackend(cfg, y)
    backend.o.set_preconditioner(cycles, left=left)
    tot = {}

    def integrate(y, tend, h, t0, rtol, atol, max_steps):
        rc, st = backend.o.integrate_adaptive(y, tend, h, t0=t0, rtol=rtol, atol=atol, max_steps=max_steps, stop_at_tend=False,
                                                 hold_step_after_failure=hold)
        for k in ("convergence_failures", "error_test_failures", "newton_iterations", "linear_iterations"):
            tot[k] = tot.get(k, 0) + int(st[k])
        return rc, st

    backend.integrate = integrate
    t0 = time.time()
    try:
        cycles_done, t, hist = run_deck.run(db, cfg, y, backend, out=io.StringIO())
    finally:
        backend.close()
    frac = [h[2]["solid_fraction"] for h in hist]
    integral = [h[2]["integral_concentration"] for h in hist]
    move = max([abs(frac[i - 1] - frac[i]) for i in range(4, len(frac))] or [float("nan")])
    print("%-5s cycles %2d%s: %d steps, t = %.5f, outputs %d, largest move of the solid fraction from the fifth output %.2e, "
          "integral c %.4f ... %.4f, %s, %.0f s" % ("left" if left else "right", cycles, " hold" if hold else "", cycles_done, t, len(frac), move,
                                                    min(integral), max(integral), tot, time.time() - t0), flush=True)


if __name__ == "__main__":
    runs = sys.argv[1:] or ["right:2", "left:2", "left:10"]
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tools", "netcdf4_shim") + os.pathsep + os.environ.get("PYTHONPATH", ""))
        subprocess.check_call([sys.executable, REF + "/tests/ConservedVolume/make_square.py", "--nx", "64", "--ny", "64", "--nz", "1",
                               "-r", "12", "2d.nc"], cwd=tmp, env=env, stdout=subprocess.DEVNULL)
        deck = os.path.join(tmp, "2d.input")
        os.symlink(REF + "/tests/ConservedVolume/2d.input", deck)
        for r in runs:
            side, n, *opt = r.split(":")
            one(deck, tmp, side == "left", int(n), "hold" in opt)
