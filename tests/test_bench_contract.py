"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line
(CPU restatement of the reference on the host cores), and the product arm refuses to run without a
CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_line():
    # OMP_NUM_THREADS=1 is what torchrun exports for its workers: the arm must still use the host's cores
    p = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", "pfhub1a"], {"OMP_NUM_THREADS": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GCUPS" and d["metric"] == "RHS cell-updates/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "sample" in cb
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count()
    assert cb["cores"] == ncores
    assert d["config"]["workload"].startswith("pfhub1a")


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--steps", "1", "--gpus", "2"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return  # on the GPU box the product arm runs (tests -m gpu, bench)
    p = _run(["--steps", "1", "--warmup", "1", "--workload", "pfhub1a"])
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
