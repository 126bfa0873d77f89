"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/nsdp_b200.h declares (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from nsdp_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "nsdp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nsdp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/nsdp_b200.h but not exported"
    # and the Python binding types exactly the declared set
    assert sorted(_lib.SIGNATURES) == names


def test_status_strings_and_arch():
    L = _lib.lib()
    assert L.nsdp_strerror(0) == b"ok"
    assert b"invalid" in L.nsdp_strerror(-1)
    assert L.nsdp_build_arch() == b"sm_100a"
    assert L.nsdp_version().startswith(b"nsdp_b200")


def test_argument_validation_happens_before_any_cuda_call():
    L = _lib.lib()
    assert L.nsdp_fps_f32(None, 1, 8, 4, None, None) == -1
    assert L.nsdp_knn_f32(None, None, 1, 8, 8, 3, None, None, None, 0, None) == -1
    assert L.nsdp_knn_workspace_bytes(8, 4096, 4096, 10) > 0      # split scan needs a workspace
    assert L.nsdp_knn_workspace_bytes(8, 50000, 100, 7) == 0      # decoder-style call does not


def test_sass_contains_only_sm100a_code():
    out = os.popen(f"cuobjdump -lelf {_lib.LIB_PATH} 2>/dev/null").read()
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_ops_fail_loudly_on_cpu_tensors():
    """The product path has no CPU fallback: CPU tensors raise, like the reference extension
    (sampling.cpp:82-84 'CPU not supported')."""
    with pytest.raises(RuntimeError, match="CPU not supported"):
        ops.furthest_point_sampling(torch.rand(1, 16, 3), 4)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        ops.knn(torch.rand(1, 16, 3), torch.rand(1, 16, 3), 4)


def test_missing_library_is_an_error(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnsdp_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or eager fallback"):
        _lib.lib()
