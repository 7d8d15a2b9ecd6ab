"""CPU: the C-ABI library loads without a GPU driver and exports every symbol include/xkv_b200.h declares."""
import os
import re

from xkv_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "xkv_b200.h")).read()
    return sorted(set(re.findall(r"XKV_API\s+[\w\s\*]+?\b(xkv_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/xkv_b200.h but not exported"


def test_bindings_cover_the_header():
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_version_and_error_channel():
    lib = _lib.load()
    assert lib.xkv_version() >= 100
    assert isinstance(lib.xkv_last_error(), bytes)
    assert lib.xkv_launch_count() >= 0
    # pure host arithmetic works without a device
    assert lib.xkv_factorize_workspace_bytes(1, 4096, 4096, 512, None) > 0
    assert lib.xkv_factorize_workspace_bytes(1, 64, 4096, 512, None) == 0      # rank > tokens
    assert b"rank" in lib.xkv_last_error()
    assert lib.xkv_decode_workspace_bytes(32, 65536, 16, 768) > 0
    # the ctypes mirrors have the library's struct layout
    import ctypes as C

    assert lib.xkv_factorize_options_size() == C.sizeof(_lib.FactorizeOptions)


def test_python_option_defaults_are_the_library_defaults():
    """FactorizeOptions() on the Python side and xkv_factorize_default_options() must describe the same factorisation: a
    default changed on one side only (solve_terms, power_terms, shifts ...) would make bench.py and the C-ABI callers of
    INTEGRATION.md run different algorithms."""
    import ctypes as C

    import pytest

    from xkv_b200 import factorize

    lib = _lib.load()
    c_def = _lib.FactorizeOptions()
    lib.xkv_factorize_default_options(C.byref(c_def))
    py_def = factorize._c_options(factorize.FactorizeOptions())
    for name, ctype in _lib.FactorizeOptions._fields_:
        a, b = getattr(c_def, name), getattr(py_def, name)
        if hasattr(a, "__len__"):
            assert list(a) == pytest.approx(list(b)), name
        elif isinstance(a, float):
            assert a == pytest.approx(b), name
        else:
            assert a == b, name
