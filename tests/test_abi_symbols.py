"""CPU: the C-ABI shared library loads and exports exactly what include/dwb.h declares (no compute calls)."""
import ctypes
import os

import pytest

from distil_whisper_b200 import _abi


def test_header_and_binding_agree():
    hs = set(_abi.header_symbols())
    assert hs == set(_abi.SIGNATURES), (sorted(hs ^ set(_abi.SIGNATURES)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(_abi.LIB_PATH)
    missing = [s for s in _abi.header_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib2 = _abi.load()
    assert _abi.call("dwb_abi_version") == 2
    assert isinstance(lib2.dwb_last_error(), bytes)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_abi.DwbError):
        _abi.call("dwb_check_device")


def test_product_package_never_imports_the_oracle():
    root = os.path.dirname(_abi.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src, f
