"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/aline_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from aline_b200 import build
    return build.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "aline_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aline_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for s in ("aline_spce_step", "aline_spce_history", "aline_lse_combine", "aline_log_likelihood",
              "aline_abi_version", "aline_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/aline_b200.h but not exported"
    L.aline_abi_version.restype = ctypes.c_int
    assert L.aline_abi_version() >= 1


def test_scratch_query_is_pure_host(built_lib):
    L = ctypes.CDLL(built_lib)
    L.aline_spce_scratch_bytes.restype = ctypes.c_size_t
    L.aline_spce_scratch_bytes.argtypes = [ctypes.c_int32, ctypes.c_int32]
    assert L.aline_spce_scratch_bytes(0, 5) == 0
    a, b = L.aline_spce_scratch_bytes(200, 35), L.aline_spce_scratch_bytes(200, 1)
    assert a > b > 0


def test_binding_signatures_load(built_lib):
    from aline_b200 import _lib
    assert _lib.lib().aline_abi_version() >= 1


def test_no_cpu_fallback():
    """A host tensor must be refused, not silently computed on the CPU."""
    import torch
    from aline_b200 import spce, AlineError
    from aline_b200.tasks import HiddenLocation
    task = HiddenLocation()
    with pytest.raises(AlineError):
        spce.spce_history(task.log_likelihood, torch.zeros(2, 3, 1), torch.zeros(2, 3, 2), torch.zeros(5, 2, 1, 2))
    with pytest.raises(AlineError):
        spce.lik_of(lambda y, xi, th: y)       # unknown log_prob callable
