"""CPU tier: the C-ABI library loads and exports every symbol include/bo_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from bayesianoptimizer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in bo_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"


def test_abi_version_and_struct_layout():
    lib = _lib.load()
    assert lib.bo_abi_version() == 1
    assert ctypes.sizeof(_lib.BoSobol) == 4 + 16 * 30 * 4 + 16 * 4


def test_no_device_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.bo_device_count() == 0
    h = ctypes.c_void_p()
    assert lib.bo_create(ctypes.byref(h), 0) == _lib.E_CUDA
    from bayesianoptimizer_b200 import BoLibraryError, GPEngine
    with pytest.raises(BoLibraryError):
        GPEngine()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bayesianoptimizer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
