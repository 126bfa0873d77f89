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


def test_fused_mlp_entry_points_validate_on_the_host():
    """nsdp_fused_mlp_*: argument checks and workspace queries are pure host code (no GPU needed)."""
    L = _lib.lib()
    a = _lib.MlpArgs()
    assert L.nsdp_fused_mlp_fwd_f32(ctypes.byref(a), None, None, 0, None) == -1           # null pointers
    for f in ("x", "w_in_t", "b_in", "w_h_t", "b_h", "w_out_t", "b_out"):
        setattr(a, f, 0x1000)                                                               # never dereferenced on the host
    a.R, a.Cin, a.W, a.O, a.n_hidden, a.impl = 1000, 3, 256, 3, 6, 0
    fwd = L.nsdp_fused_mlp_fwd_workspace_bytes(ctypes.byref(a))
    assert fwd == 6 * 16 * 16384 + 16                                                       # 6 layers x 16 k-steps x (hi + lo slab)
    bwd = L.nsdp_fused_mlp_bwd_workspace_bytes(ctypes.byref(a))
    assert bwd == 2 * (fwd - 16) + 256 + 8 * 512 * (2 * 7 * 256 + 32)                       # both weight images + 8 staged tiles
    a.W, a.impl = 100, 2
    assert L.nsdp_fused_mlp_fwd_workspace_bytes(ctypes.byref(a)) == 0
    assert L.nsdp_fused_mlp_fwd_f32(ctypes.byref(a), 0x1000, None, 0, None) == -2           # tcgen05 required, width not instantiated
    assert L.nsdp_fused_mlp_bwd_f32(ctypes.byref(a), 0x1000, None, None, 0, None) == -1     # no gradient struct
    a.W, a.Cin = 256, 5
    assert L.nsdp_fused_mlp_fwd_f32(ctypes.byref(a), 0x1000, None, 0, None) == -2           # Cin > 4


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
    w = [torch.rand(16, 3), torch.rand(16), torch.rand(1, 16, 16), torch.rand(1, 16), torch.rand(3, 16), torch.rand(3)]
    with pytest.raises(RuntimeError, match="CPU not supported"):
        ops.FusedMLP(*w)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        ops.fused_mlp(torch.rand(8, 3), w[0].t().contiguous(), w[1], w[2], w[3], w[4].t().contiguous(), w[5])


def test_missing_library_is_an_error(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnsdp_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or eager fallback"):
        _lib.lib()
