"""Host accuracy check of the straight-line atan prepared for the bias-well term (ampe_b200/csrc/atan_core.h):
the same fma chain compiled with g++ against atanl over 1e-12 .. 1e12 and around the range boundaries."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_atan_fast_core_accuracy(tmp_path):
    exe = str(tmp_path / "atan_accuracy")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "atan_accuracy.cpp")])
    out = dict(line.split() for line in subprocess.check_output([exe]).decode().strip().splitlines())
    assert float(out["max_ulp"]) < 2.5
    assert float(out["max_rel"]) < 5e-16  # four orders below the 1e-12 parity bar of the phase RHS
    assert float(out["atan0"]) == 0.0
