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
