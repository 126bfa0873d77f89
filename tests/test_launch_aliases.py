"""The unchanged reference scripts import `model` / `pointnet2_ops_lib...`: nsdp_b200.launch must bind those names to
the mirrors (CPU-only check of the import plumbing)."""
import importlib
import sys

import pytest


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


def test_per_rank_seed_and_output_directory(tmp_path):
    """Under torchrun every rank of an unchanged train.py gets a distinct --seed, and ranks > 0 a private out_dir
    (design A)."""
    import yaml
    from nsdp_b200 import launch
    cfg = tmp_path / "forward.yaml"
    cfg.write_text(yaml.safe_dump({"experiment": {"out_dir": "/data/out", "name": "run"}, "model": {"type": "forward"}}))
    cfg.write_text(yaml.safe_dump({"experiment": {"out_dir": str(tmp_path / "out"), "name": "run"}, "model": {"type": "forward"}}))
    argv = ["/repo/NSDP/train.py", str(cfg), "--num_workers", "4"]
    assert launch.rewrite_argv_for_rank(argv, 0, 1, str(tmp_path)) == argv                       # single process: untouched
    r0 = launch.rewrite_argv_for_rank(argv, 0, 8, str(tmp_path / "s"))
    assert r0 == argv + ["--seed", "27"]                                                          # rank 0 keeps the config
    r3 = launch.rewrite_argv_for_rank(argv + ["--seed", "100"], 3, 8, str(tmp_path / "s"))
    assert r3[-2:] == ["--seed", "103"] and r3.count("--seed") == 1 and r3[2:4] == ["--num_workers", "4"]
    got = yaml.safe_load(open(r3[1]))
    assert got["experiment"] == {"out_dir": str(tmp_path / "out" / "rank3"), "name": "run"} and got["model"] == {"type": "forward"}
    test_argv = ["/repo/NSDP/test.py", str(cfg)]
    assert launch.rewrite_argv_for_rank(test_argv, 3, 8, str(tmp_path)) == test_argv              # only train.py writes checkpoints


def test_resume_reads_rank0_checkpoints_on_every_rank(tmp_path):
    """ranks > 0 must not resume from their own (stale / missing) files: their private experiment directory
    starts as links to rank 0's model_* / opt_* / modelbest_* files, so load_checkpoints (utils/checkpoints.py:8-31)
    picks the same epoch and weights everywhere."""
    import os
    import yaml
    from nsdp_b200 import launch
    shared = tmp_path / "out" / "run"
    shared.mkdir(parents=True)
    for f in ("model_00020", "opt_00020", "modelbest_00010_0.123000", "stats.txt"):
        (shared / f).write_text(f)
    stale = tmp_path / "out" / "rank1" / "run"
    stale.mkdir(parents=True)
    (stale / "model_00040").write_text("from an older 2-rank run")
    (stale / "opt_00040").write_text("x")
    cfg = tmp_path / "forward.yaml"
    cfg.write_text(yaml.safe_dump({"experiment": {"out_dir": str(tmp_path / "out"), "name": "run"}}))
    launch.rewrite_argv_for_rank(["/repo/NSDP/train.py", str(cfg)], 1, 2, str(tmp_path / "s"))
    names = sorted(os.listdir(stale))
    assert names == ["model_00020", "modelbest_00010_0.123000", "opt_00020"]
    assert (stale / "model_00020").read_text() == "model_00020" and os.path.islink(stale / "opt_00020")


def test_rank_device_is_taken_from_the_devices_the_job_was_given():
    from nsdp_b200.launch import device_for_local_rank
    assert device_for_local_rank(3, None) == "3" and device_for_local_rank(0, "") == "0"
    assert device_for_local_rank(1, "4,5,6,7") == "5"                       # a job confined to the second half of the box
    assert device_for_local_rank(0, "GPU-aaaa, GPU-bbbb") == "GPU-aaaa"
    with pytest.raises(SystemExit):
        device_for_local_rank(2, "0,1")
