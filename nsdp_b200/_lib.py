"""ctypes binding of libnsdp_b200.so (include/nsdp_b200.h).

This is the only place the shared library is opened. There is NO fallback: if the library is missing
or a call fails, a RuntimeError is raised — the product path never routes through torch eager or the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NSDP_B200_LIB") or os.path.join(_PKG, "lib", "libnsdp_b200.so")  # override: A/B builds

c_float_p = C.c_void_p
c_int_p = C.c_void_p


class VattnArgs(C.Structure):
    _fields_ = [
        ("xyz_c", C.c_void_p), ("xyz_n", C.c_void_p), ("idx", C.c_void_p),
        ("qp", C.c_void_p), ("kp", C.c_void_p), ("vp", C.c_void_p),
        ("gq", C.c_void_p), ("gv", C.c_void_p),
        ("wd0", C.c_void_p), ("bd0", C.c_void_p), ("wd2t", C.c_void_p), ("wpt", C.c_void_p),
        ("wg2t", C.c_void_p), ("pc", C.c_void_p), ("vc", C.c_void_p),
        ("wd2", C.c_void_p), ("wp", C.c_void_p), ("wg2", C.c_void_p),
        ("B", C.c_int), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("D", C.c_int),
        ("has_global", C.c_int), ("sign", C.c_float), ("impl", C.c_int),
        ("saved", C.c_void_p), ("saved_bytes", C.c_size_t),
    ]


class VattnGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "d_qp", "d_kp", "d_vp", "d_gq", "d_gv", "d_wd0", "d_bd0", "d_wd2t", "d_wpt", "d_wg2t", "d_pc", "d_vc",
        "d_xyz_c", "d_xyz_n")]


class TailArgs(C.Structure):
    _fields_ = [
        ("lat", C.c_void_p), ("wc_t", C.c_void_p), ("bc", C.c_void_p),
        ("w0_t", C.c_void_p), ("b0", C.c_void_p), ("w1_t", C.c_void_p), ("b1", C.c_void_p),
        ("wo_t", C.c_void_p), ("bo", C.c_void_p),
        ("R", C.c_int), ("C", C.c_int), ("H", C.c_int), ("O", C.c_int), ("n_blocks", C.c_int), ("impl", C.c_int),
    ]


class TailGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "d_lat", "d_wc_t", "d_bc", "d_w0_t", "d_b0", "d_w1_t", "d_b1", "d_wo_t", "d_bo")]


class MlpArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w_in_t", C.c_void_p), ("b_in", C.c_void_p),
        ("w_h_t", C.c_void_p), ("b_h", C.c_void_p), ("w_out_t", C.c_void_p), ("b_out", C.c_void_p),
        ("R", C.c_int), ("Cin", C.c_int), ("W", C.c_int), ("O", C.c_int), ("n_hidden", C.c_int),
        ("impl", C.c_int), ("reuse_packed", C.c_int),
    ]


class MlpGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_x", "d_w_in_t", "d_b_in", "d_w_h_t", "d_b_h", "d_w_out_t", "d_b_out")]


class EmlpArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("bn_weight", C.c_void_p * 3), ("bn_bias", C.c_void_p * 3),
        ("running_mean", C.c_void_p * 3), ("running_var", C.c_void_p * 3), ("num_batches_tracked", C.c_void_p * 3),
        ("R", C.c_int), ("C", C.c_int), ("training", C.c_int), ("momentum", C.c_float), ("eps", C.c_float),
    ]


class EmlpGrads(C.Structure):
    _fields_ = [("d_x", C.c_void_p), ("d_w1", C.c_void_p), ("d_b1", C.c_void_p), ("d_w2", C.c_void_p), ("d_b2", C.c_void_p),
                ("d_bn_weight", C.c_void_p * 3), ("d_bn_bias", C.c_void_p * 3)]


# name -> (restype, argtypes); must list every symbol include/nsdp_b200.h declares
# (tests/test_abi.py cross-checks this table against the header).
_P, _I, _F, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_size_t
SIGNATURES = {
    "nsdp_strerror": (C.c_char_p, [_I]),
    "nsdp_last_cuda_error": (_I, []),
    "nsdp_version": (C.c_char_p, []),
    "nsdp_build_arch": (C.c_char_p, []),
    "nsdp_fps_f32": (_I, [_P, _I, _I, _I, _P, _P]),
    "nsdp_gather_points_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "nsdp_gather_points_grad_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "nsdp_ball_query_f32": (_I, [_P, _P, _I, _I, _I, _F, _I, _P, _P]),
    "nsdp_group_points_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nsdp_group_points_grad_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "nsdp_three_nn_f32": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "nsdp_three_interpolate_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "nsdp_three_interpolate_grad_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "nsdp_knn_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "nsdp_knn_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "nsdp_vattn_fwd_workspace_bytes": (_SZ, [C.POINTER(VattnArgs)]),
    "nsdp_vattn_saved_bytes": (_SZ, [C.POINTER(VattnArgs)]),
    "nsdp_vattn_fwd_f32": (_I, [C.POINTER(VattnArgs), _P, _P, _P, _SZ, _P]),
    "nsdp_vattn_bwd_workspace_bytes": (_SZ, [C.POINTER(VattnArgs)]),
    "nsdp_vattn_bwd_f32": (_I, [C.POINTER(VattnArgs), _P, _P, _P, C.POINTER(VattnGrads), _P, _SZ, _P]),
    "nsdp_resnet_tail_fwd_workspace_bytes": (_SZ, [C.POINTER(TailArgs)]),
    "nsdp_resnet_tail_fwd_f32": (_I, [C.POINTER(TailArgs), _P, _P, _SZ, _P]),
    "nsdp_resnet_tail_bwd_workspace_bytes": (_SZ, [C.POINTER(TailArgs)]),
    "nsdp_resnet_tail_bwd_f32": (_I, [C.POINTER(TailArgs), _P, C.POINTER(TailGrads), _P, _SZ, _P]),
    "nsdp_fused_mlp_fwd_workspace_bytes": (_SZ, [C.POINTER(MlpArgs)]),
    "nsdp_fused_mlp_fwd_f32": (_I, [C.POINTER(MlpArgs), _P, _P, _SZ, _P]),
    "nsdp_fused_mlp_bwd_workspace_bytes": (_SZ, [C.POINTER(MlpArgs)]),
    "nsdp_fused_mlp_bwd_f32": (_I, [C.POINTER(MlpArgs), _P, C.POINTER(MlpGrads), _P, _SZ, _P]),
    "nsdp_emlp_stats_bytes": (_SZ, [C.POINTER(EmlpArgs)]),
    "nsdp_emlp_fwd_f32": (_I, [C.POINTER(EmlpArgs), _P, _P, _P, _P, _P, _P]),
    "nsdp_emlp_bwd_workspace_bytes": (_SZ, [C.POINTER(EmlpArgs)]),
    "nsdp_emlp_bwd_f32": (_I, [C.POINTER(EmlpArgs), _P, _P, _P, _P, _P, C.POINTER(EmlpGrads), _P, _SZ, _P]),
    "nsdp_set_stage_format": (_I, [_I]),
    "nsdp_linear_narrow_dw_f32": (_I, [_P, _P, C.c_longlong, _I, _I, _P, _P, _P]),
    "nsdp_adam_step_f32": (_I, [_I, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _P]),
    "nsdp_selftest_umma": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "nsdp_selftest_umma2": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Open libnsdp_b200.so (once) and type every entry point. Raises if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m nsdp_b200.build` "
                "(nsdp_b200 has no CPU or eager fallback).")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        l = lib()
        msg = l.nsdp_strerror(status).decode()
        extra = ""
        if status == -3:
            extra = f" (cudaError_t={l.nsdp_last_cuda_error()})"
        raise RuntimeError(f"{what} failed: {msg}{extra}")
