"""Imports the unmodified reference `model` package from baseline/_ref/ (see make_ref.py).

The reference imports `pointnet2_ops_lib.pointnet2_ops.pointnet2_utils` (model/encoder/blocks.py:15), which needs a
top-level `pointnet2_ops` package with a compiled `_ext` (pointnet2_ops/__init__.py:1-3, pointnet2_utils.py:8). The
compiled module is supplied here, never by the product:
  * CUDA tensors -> the reference's OWN extension, built from the unmodified sources by oracle/build_ref.py
    (oracle/_ref/nsdp_ref_pointnet2_ext.so);
  * CPU tensors  -> the reference has no CPU kernel ("CPU not supported", sampling.cpp:82-84); furthest_point_sampling
    falls back to the C restatement oracle/nsdp_oracle.c (pinned index-for-index against the real kernel on the GPU
    box, tests/test_gpu_index_kernels.py). The other eight operators are never called by the model.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
_MODEL = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "model")) and os.path.isdir(os.path.join(REF, "pointnet2_ops_lib"))


def _make_ext():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_ext
    ext = types.ModuleType("pointnet2_ops._ext")
    state = {"cuda": None}

    def cuda_ext():
        if state["cuda"] is None:
            state["cuda"] = ref_ext.load()
            if state["cuda"] is None:
                raise RuntimeError("oracle/_ref/nsdp_ref_pointnet2_ext.so is missing: run oracle/build_ref.py where /root/reference exists")
        return state["cuda"]

    def furthest_point_sampling(xyz, npoint):
        if xyz.is_cuda:
            return cuda_ext().furthest_point_sampling(xyz, npoint)
        from oracle import tdnet_oracle as orc
        return orc.fps(xyz, npoint)

    ext.furthest_point_sampling = furthest_point_sampling
    for name in ("gather_points", "gather_points_grad", "ball_query", "group_points", "group_points_grad", "three_nn",
                 "three_interpolate", "three_interpolate_grad"):
        ext.__dict__[name] = (lambda n: lambda *a: getattr(cuda_ext(), n)(*a))(name)
    return ext


def load():
    """Returns the reference's `model` package (build_model, optimizer_factory, ...)."""
    global _MODEL
    if _MODEL is not None:
        return _MODEL
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python baseline/make_ref.py` where /root/reference exists")
    for name in ("model", "pointnet2_ops", "pointnet2_ops_lib"):
        if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(REF):
            raise RuntimeError(f"a different '{name}' module is already imported (nsdp_b200.launch aliases?)")
    sys.modules["pointnet2_ops._ext"] = _make_ext()
    sys.path.insert(0, os.path.join(REF, "pointnet2_ops_lib"))
    sys.path.insert(0, REF)
    try:
        import model as ref_model
    finally:
        sys.path.remove(REF)
    _MODEL = ref_model
    return ref_model
