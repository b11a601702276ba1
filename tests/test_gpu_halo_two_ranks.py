"""Slab ranks as separate processes on the GPU box (one GPU is enough: CUDA IPC maps memory between processes on
the same device): the C-ABI ghost-plane exchange (csrc/halo.cu) against the one-rank evaluation of the whole
periodic domain, bit for bit.

* tests/cpp/halo_two_ranks.cpp: two forked processes, no Python / MPI / NCCL; handles over a socketpair.
* tools/mgpu_check.py under torchrun with two ranks: every workload family incl. the symmetry-aware path, the
  Newton reference concentrations, fd_flag 0 / 1, the host-buffer path, with and without overlap."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_cpp_two_processes(tmp_path):
    exe = str(tmp_path / "halo_two_ranks")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", cuda + "/include",
                           os.path.join(ROOT, "tests", "cpp", "halo_two_ranks.cpp"), "-L", os.path.join(ROOT, "ampe_b200"),
                           "-lampe_b200", "-L", cuda + "/lib64", "-lcudart",
                           "-Wl,-rpath," + os.path.join(ROOT, "ampe_b200"), "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "HALO TWO RANKS OK" in p.stdout


@pytest.mark.parametrize("world", [2, 3])
def test_torchrun_ranks_share_or_split_the_gpus(world):
    env = dict(os.environ)
    env.pop("AMPE_B200_HALO", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "mgpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
    assert "MGPU CHECK OK" in p.stdout


@pytest.mark.timeout(1200)
def test_regression_decks_on_slab_ranks():
    """tools/mgpu_deck.py: the Dendrite deck integrated in full by the variable-step implicit integrator on two slab ranks
    (zero-slope boundaries: the exchange ring is cut; vector reductions through the sum-reduction hook) lands on the
    one-rank solid fraction and on the deck's acceptance number.  AMPE_B200_SLOW_DECKS=1 adds the AuNi decks
    (unpreconditioned start; full run with block-Jacobi preconditioners) and TwoGrainsQuadratic 3D; all four cases on
    2, 4 and 8 ranks: profiles/r02t_mgpu_deck_n{2,4}.log, r02u_mgpu_deck_n2.log, r02y_mgpu_deck_n8.log (two processes
    time-slicing one GPU are slow)."""
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tools", "mgpu_deck.py")] + (
               [] if os.environ.get("AMPE_B200_SLOW_DECKS") else ["0"])
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=1100, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
    assert "MGPU DECK OK" in p.stdout
