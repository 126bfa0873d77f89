"""TEST INFRASTRUCTURE — loads the REFERENCE's pointnet2_ops CUDA extension built by oracle/build_ref.py
(from the unmodified sources under /root/reference, for sm_100a) so that `-m gpu` tests can compare the
product kernels and oracle/nsdp_oracle.c against the real thing on the B200 box. Returns None when the
binary is absent (it can only be built where /root/reference exists)."""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "nsdp_ref_pointnet2_ext.so")
_MOD = None


def load():
    global _MOD
    if _MOD is None:
        if not os.path.exists(_PATH):
            return None
        import torch  # noqa: F401  (libtorch symbols must be loaded first)
        spec = importlib.util.spec_from_file_location("nsdp_ref_pointnet2_ext", _PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _MOD = mod
    return _MOD
