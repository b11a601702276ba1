#!/usr/bin/env python
"""Find explicit-Euler step sizes for the trajectory parity tests with the CPU oracle:
largest dt (decade search) for which 60 steps keep max|ydot| from growing (test tooling)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from oracle import pyoracle

for name in (sys.argv[1:] or list(parity.SMALL)):
    cfg, st = parity.make_case(name)
    rot = parity.random_rotations(cfg) if cfg.symmetry_aware else None
    y0 = {k: (None if v is None else v.numpy().copy()) for k, v in st.items()}
    o = pyoracle.Oracle(cfg)
    if cfg.conc_rhs_form in (2, 3):
        o.set_ref(y0["conc"].ravel().copy(), y0["conc"].ravel().copy())
    if rot: o.set_rotations(rot)
    _, yd = o.eval(0.0, y0, 0)
    amp = {k: (np.abs(y0[k]).max(), np.abs(v).max()) for k, v in yd.items() if v is not None}
    print(name, "max|y|, max|ydot|:", {k: "%.2e %.2e" % a for k, a in amp.items()})
    o.close()
    for e in range(0, -12, -1):
        for mant in (5.0, 2.0, 1.0):
            dt = mant * 10.0 ** e
            try:
                y, _ = parity.oracle_trajectory(cfg, st, dt, 60, rot)
                o = pyoracle.Oracle(cfg)
                if cfg.conc_rhs_form in (2, 3):
                    o.set_ref(y["conc"].ravel().copy(), y["conc"].ravel().copy())
                if rot: o.set_rotations(rot)
                stt, yd2 = o.eval(0.0, y, 0)
                o.close()
                ok = stt == 0 and all(np.isfinite(v).all() and np.abs(v).max() <= 1.5 * amp[k][1] + 1e-300
                                      for k, v in yd2.items() if v is not None)
            except AssertionError:
                ok = False
            if ok:
                break
        if ok:
            print("  stable dt ~", dt)
            break
