#!/usr/bin/env python
"""Generate the initial conditions of the reference's regression decks by running the reference's OWN generator,
utils/make_nuclei.py, unmodified, with the command lines its tests use (tests/Dendrite/test2d.py:11-15,
tests/SingleGrainGrowthAuNi/test2d.py:11-15, tests/TwoGrainsQuadratic/test3d.py:11-15).  The generator writes
NetCDF-4 through the `netCDF4` module, which this image does not have: a stand-in module with the handful of calls
the script makes (Dataset / createDimension / createVariable / var[:, :, :] = array / close) captures the arrays,
which are committed as compressed fixtures under tests/golden/ (single precision, as the generator writes them).
Run in the build container only (needs /root/reference):   python tools/make_reference_nuclei.py"""
import os
import runpy
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_UTILS = "/root/reference/utils"

DECKS = {
    # name: argv of make_nuclei.py as in the reference's test script
    "dendrite": ["--nx", "240", "--ny", "240", "--nz", "1", "-r", "16", "--center0", "0, 0, 0", "-w", "1.4"],
    "single_grain_auni": ["--nx", "64", "--ny", "64", "--nz", "1", "-r", "16", "--center0", "0, 0, 0", "-c", "0.25",
                          "--concentration-in", "0.096"],
    "two_grains_quadratic": ["--nx", "64", "--ny", "64", "--nz", "48", "-r", "8", "--concentration-in", "0.1",
                             "--concentration-out", "0.06", "--ngrains", "2", "-q", "4"],
    # the 3D versions of the decks: tests/Dendrite/test3d.py:11-15, tests/SingleGrainGrowthAuNi/test3d.py:10-14 (=
    # tests/KKScomposition/test3d.py:10-14), tests/FourCorners/test3d.py:11-13
    "dendrite3d": ["--nx", "60", "--ny", "60", "--nz", "60", "-r", "10", "--center0", "0, 0, 0", "-w", "1.4"],
    "single_grain_auni3d": ["--nx", "32", "--ny", "32", "--nz", "32", "-r", "16", "--center0", "0, 0, 0", "-c", "0.25",
                            "--concentration-in", "0.096"],
    "four_corners3d": ("make4corners.py", ["-x", "64", "-y", "64", "-z", "4"]),
    # tests/OneGrainQuadratic/test2d.py:11-15, test3d.py:11-15
    "one_grain_quadratic2d": ["--nx", "64", "--ny", "64", "--nz", "1", "-r", "8", "--concentration-in", "0.1",
                              "--concentration-out", "0.06", "--ngrains", "1"],
    "one_grain_quadratic3d": ["--nx", "48", "--ny", "48", "--nz", "48", "-r", "8", "--concentration-in", "0.1",
                              "--concentration-out", "0.06", "--ngrains", "1"],
    # tests/TwoGrainsQuadratic/test2d.py:11-15
    "two_grains_quadratic2d": ["--nx", "64", "--ny", "64", "--nz", "1", "-r", "8", "--concentration-in", "0.1",
                               "--concentration-out", "0.06", "--ngrains", "2", "-q", "4"],
    # tests/FourCorners/test2d.py:11-13 (utils/make4corners.py), tests/SolidifyQuaternions/test2d.py:11-14
    # (utils/make_initial_grains_on_boundary.py, random.seed(112345) inside)
    "four_corners": ("make4corners.py", ["-x", "64", "-y", "64", "-z", "1"]),
    # tests/SolidifyQuaternions/test3d.py:11-14
    "solidify_quaternions3d": ("make_initial_grains_on_boundary.py",
                               ["-x", "48", "-y", "24", "-z", "8", "--solid-fraction", "0.25", "--smooth", "0", "--qlen", "4",
                                "--ngrains", "2"]),
    "solidify_quaternions": ("make_initial_grains_on_boundary.py",
                             ["-x", "64", "-y", "32", "-z", "1", "--solid-fraction", "0.25", "--smooth", "0", "--qlen", "4",
                              "--ngrains", "2"]),
}


class _Var:
    def __init__(self, store, name, dtype, shape):
        self.store, self.name = store, name
        self.store[name] = np.zeros(shape, dtype=np.float32 if dtype == "f" else np.float64)

    def __setitem__(self, key, value):
        self.store[self.name][key] = value


class _Dataset:
    captured = None

    def __init__(self, filename, mode="w", format=None):
        self.dims, self.vars = {}, {}
        _Dataset.captured = self.vars

    def createDimension(self, name, n):
        self.dims[name] = n

    def createVariable(self, name, dtype, dims):
        return _Var(self.vars, name, dtype, tuple(self.dims[d] for d in dims))

    def close(self):
        pass


def run(argv, script="make_nuclei.py"):
    fake = types.ModuleType("netCDF4")
    fake.Dataset = _Dataset
    sys.modules["netCDF4"] = fake
    sys.path.insert(0, REF_UTILS)
    old = sys.argv
    sys.argv = [script] + argv + ["unused.nc"]
    # make_initial_grains_on_boundary.py calls random.randint with integral FLOAT bounds, which Python <= 3.9 accepted
    # (deprecated in 3.10, a TypeError since 3.12): give it the interpreter it was written for -- same Mersenne Twister,
    # same draws
    import random
    _randint = random.randint
    random.randint = lambda a, b: _randint(int(a), int(b))
    try:
        runpy.run_path(os.path.join(REF_UTILS, script), run_name="__main__")
    finally:
        random.randint = _randint
        sys.argv = old
        sys.path.remove(REF_UTILS)
    return dict(_Dataset.captured)


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    only = sys.argv[1:]
    for name, argv in DECKS.items():
        if only and name not in only:
            continue
        fields = run(*((argv[1], argv[0]) if isinstance(argv, tuple) else (argv,)))
        path = os.path.join(out, "ic_%s.npz" % name)
        np.savez_compressed(path, **fields)
        print(name, {k: (v.shape, str(v.dtype), float(v.min()), float(v.max())) for k, v in fields.items()},
              "%d bytes" % os.path.getsize(path))
