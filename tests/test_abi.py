"""The C-ABI library loads without a GPU, exports every symbol that include/*.h
declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from ampe_b200 import _abi, configs, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if fn.endswith(".h"):
            text = open(os.path.join(inc, fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names += re.findall(r"\b(ampe_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    L = lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_struct_layout_matches_library():
    L = lib.load()
    assert L.ampe_abi_sizeof_config() == C.sizeof(_abi.RhsConfig)
    assert b"sm_100a" in L.ampe_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = lib.load()
    cfg = configs.dendrite2d(nx=64, ny=64)
    h = C.c_void_p()
    rc = L.ampe_rhs_create(C.byref(cfg), C.byref(h))
    assert rc == _abi.AMPE_ENOGPU
    assert b"no CPU fallback" in L.ampe_last_error()
    from ampe_b200 import rhs
    with pytest.raises(lib.AmpeError):
        rhs.QuatIntegratorRHS(cfg)


def test_product_never_imports_oracle():
    """only tests/, smoke() and bench.py's baseline legs may touch oracle/"""
    pkg = os.path.join(ROOT, "ampe_b200")
    for base, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                text = open(os.path.join(base, fn)).read()
                assert "oracle" not in text.lower() or fn == "__init__.py", os.path.join(base, fn)
