"""The C restatement of the index kernels (oracle/nsdp_oracle.c) held to vectors minted by the LIVE reference's own torch code
(tests/golden/make_golden_index.py: farthest_point_sample with start 0, square_distance + argsort, index_points — all imported
unmodified from /root/reference/model/utils.py). No GPU. The CUDA kernels are held to the same vectors in
tests/test_gpu_index_kernels.py."""
import os

import numpy as np
import pytest
import torch

from oracle import tdnet_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "index_reference.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def cases(gold, kind):
    return sorted({k.split("::")[1] for k in gold.keys() if k.startswith(kind + "::")})


def test_fixture_has_every_case(gold):
    assert cases(gold, "fps") == ["all_points", "bumpy_fp16_4096", "bumpy_fp16_500", "bumpy_fp32_2048", "tiny_37", "uniform_1000"]
    assert cases(gold, "knn") == ["k_equals_n", "q300_r1000_k16", "q500_r100_k7", "self_800_k10"]


def test_c_oracle_fps_equals_reference_torch_fps(gold):
    for name in cases(gold, "fps"):
        xyz = torch.from_numpy(gold[f"fps::{name}::xyz"])
        want = torch.from_numpy(gold[f"fps::{name}::idx"])
        got = orc.fps(xyz, want.shape[1])
        assert got.dtype == torch.int32
        assert torch.equal(got, want), f"{name}: first mismatch at {(got != want).nonzero()[:3].tolist()}"
        assert bool((got[:, 0] == 0).all())                                     # sampling_gpu.cu:84-86: starts at index 0
        assert all(row.unique().numel() == row.numel() for row in got)          # no point picked twice


def test_c_oracle_knn_equals_reference_argsort(gold):
    for name in cases(gold, "knn"):
        q = torch.from_numpy(gold[f"knn::{name}::query"])
        r = torch.from_numpy(gold[f"knn::{name}::ref"])
        want = torch.from_numpy(gold[f"knn::{name}::idx"])
        got, d2 = orc.knn(q, r, want.shape[2], return_d2=True)
        assert torch.equal(got, want), name
        assert torch.equal(d2, torch.from_numpy(gold[f"knn::{name}::d2"])), name   # same association, no FMA: bit-identical
        assert bool((d2[:, :, 1:] >= d2[:, :, :-1]).all())                          # ascending


def test_index_points_restatements_equal_reference(gold):
    feats = torch.from_numpy(gold["index_points::feats"])
    for tag in ("2", "3"):
        idx = torch.from_numpy(gold[f"index_points::idx{tag}"])
        want = torch.from_numpy(gold[f"index_points::out{tag}"])
        assert torch.equal(orc.index_points(feats, idx.long()), want)
        from nsdp_b200.model.utils import index_points          # the model mirror's gather (plain torch, runs anywhere)
        assert torch.equal(index_points(feats, idx), want)


def test_mirror_keeps_the_references_torch_fps_and_sphere_helpers(gold, monkeypatch):
    """model/utils.py:13-36, 73-93 exist in the mirror under the same names (API completeness; neither is on the hot path)."""
    from nsdp_b200.model import utils as mu
    monkeypatch.setattr(torch, "randint", lambda lo, hi, size, **kw: torch.zeros(size, dtype=kw.get("dtype", torch.long)))
    for name in ("bumpy_fp16_500", "tiny_37", "uniform_1000"):
        xyz = torch.from_numpy(gold[f"fps::{name}::xyz"])
        want = torch.from_numpy(gold[f"fps::{name}::idx"])
        got = mu.farthest_point_sample(xyz, want.shape[1])
        assert got.dtype == torch.int64 and torch.equal(got, want.long()), name
    monkeypatch.undo()
    start = mu.farthest_point_sample(torch.rand(64, 50, 3), 4)[:, 0]
    assert start.unique().numel() > 1                                     # the start really is random per shape
    pts = mu.fibonacci_sphere(200)
    assert pts.shape == (200, 3) and np.allclose(np.linalg.norm(pts, axis=1), 1.0)
    assert pts[0, 1] == 1.0 and pts[-1, 1] == -1.0                        # y runs from +1 to -1
    phi = np.pi * (3.0 - np.sqrt(5.0))
    assert np.allclose(pts[7], [np.cos(7 * phi) * np.sqrt(1 - pts[7, 1] ** 2), 1 - 14 / 199, np.sin(7 * phi) * np.sqrt(1 - pts[7, 1] ** 2)])


def test_mirror_keeps_weights_init_and_clamp_gradient():
    """model/learningrate.py:50-64."""
    from nsdp_b200.model import learningrate as lr
    net = torch.nn.Sequential(torch.nn.Linear(4, 5), torch.nn.BatchNorm1d(5), torch.nn.Conv1d(5, 5, 1))
    with torch.no_grad():
        net[1].weight.fill_(3.0)
        net[1].bias.fill_(2.0)
    net.apply(lr.weights_init)
    assert float(net[0].bias.abs().max()) == 0.0 and float(net[2].bias.abs().max()) == 0.0
    assert float(net[0].weight.abs().max()) <= (6.0 / 9.0) ** 0.5 + 1e-6   # Xavier-uniform bound sqrt(6 / (fan_in + fan_out))
    assert torch.equal(net[1].weight, torch.ones(5)) and torch.equal(net[1].bias, torch.zeros(5))
    net[0](torch.randn(3, 4)).mul(100.0).sum().backward()
    lr.clamp_gradient(net, 0.25)
    assert float(net[0].weight.grad.abs().max()) <= 0.25 and float(net[0].weight.grad.abs().max()) > 0.0
