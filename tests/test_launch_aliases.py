"""The unchanged reference scripts import `model` / `pointnet2_ops_lib...`: nsdp_b200.launch must bind those names to
the mirrors (CPU-only check of the import plumbing)."""
import importlib
import sys


def test_aliases_resolve_to_mirrors():
    saved = dict(sys.modules)
    try:
        from nsdp_b200 import launch
        launch.install_aliases()
        model = importlib.import_module("model")
        assert model.build_model.__module__ == "nsdp_b200.model"
        lr = importlib.import_module("model.learningrate")
        assert hasattr(lr, "adjust_learning_rate") and hasattr(lr, "get_learning_rates") and hasattr(lr, "print_num_parameters")
        p2u = importlib.import_module("pointnet2_ops_lib.pointnet2_ops.pointnet2_utils")
        assert p2u.furthest_point_sample is not None and p2u.__name__.startswith("nsdp_b200")
        ext = importlib.import_module("pointnet2_ops._ext")
        for name in ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn", "three_interpolate",
                     "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"):
            assert callable(getattr(ext, name)), name   # bindings.cpp:6-19
    finally:
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]
