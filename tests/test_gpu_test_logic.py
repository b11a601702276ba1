"""The GPU tests written without a GPU (tests/test_gpu_widening_zzz_precond.py) are exercised on CPU stand-ins
(tools/check_gpu_test_logic.py): shapes, argument order and thresholds of the test code itself.  Says nothing
about the CUDA code."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_test_logic_runs_on_cpu_stand_ins():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_gpu_test_logic.py")], capture_output=True,
                       text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-3000:]
    assert "32 test invocations exercised" in p.stdout
