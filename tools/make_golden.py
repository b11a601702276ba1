#!/usr/bin/env python
"""Generate tests/golden/oracle_<config>.npz: outputs of the CPU oracle (parity build,
-O2 -ffp-contract=off) on the seeded small cases of tests/parity.py.  They pin the
oracle against accidental change and give the GPU tests a stored vector to compare
with.  Re-run only when the oracle is deliberately changed:  python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
for name in parity.SMALL:
    cfg, st = parity.make_case(name)
    rot = parity.random_rotations(cfg) if cfg.symmetry_aware else None
    outs, extra = parity.run_oracle(cfg, st, fd_flags=(0,), rotations=rot)
    status, yd = outs[0]
    assert status == 0
    rec = {"ydot_" + k: v for k, v in yd.items() if v is not None}
    # inputs are stored too: regenerating them with another libm could differ in the last bit
    for k, v in st.items():
        if v is not None:
            rec["in_" + k] = v.numpy()
    if rot is not None:
        for d, a in enumerate(rot):
            rec["rot%d" % d] = a
    if extra is not None:
        rec["cl"], rec["ca"] = extra
    np.savez_compressed(os.path.join(out, "oracle_%s.npz" % name), **rec)
    print(name, {k: v.shape for k, v in rec.items()})
