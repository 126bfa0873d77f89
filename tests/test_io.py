"""sample-file and checkpoint formats either side of the hot path (CPU), and the host->device stager
(GPU). Format semantics follow dataset/utils.py:8-17 and utils/checkpoints.py:8-77 of the reference."""
import os
import types

import numpy as np
import pytest
import torch

from nsdp_b200 import io as nio
from nsdp_b200 import synth
from nsdp_b200.model import build_model, optimizer_factory


def test_npz_sample_formats_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    pts = rng.uniform(-0.5, 0.5, (4096, 3)).astype(np.float32)
    nrm = rng.normal(size=(4096, 3)).astype(np.float32)
    nio.save_npz_surface_flow(tmp_path / "surface_points.npz", pts, nrm)
    nio.save_npz_space_flow(tmp_path / "flow.npz", pts[:1000])
    raw = np.load(tmp_path / "surface_points.npz")
    assert raw["points"].dtype == np.float16 and raw["normals"].dtype == np.float16    # the on-disk format is fp16
    p, n = nio.load_npz_surface_flow(tmp_path / "surface_points.npz")
    q = nio.load_npz_space_flow(tmp_path / "flow.npz")
    assert p.dtype == n.dtype == q.dtype == np.float32
    np.testing.assert_array_equal(p, pts.astype(np.float16).astype(np.float32))
    np.testing.assert_array_equal(q, pts[:1000].astype(np.float16).astype(np.float32))


def test_checkpoint_files_have_the_reference_names_and_resume(tmp_path):
    cfg = synth.make_config("forward")
    model, *_ = build_model(cfg)
    _, opt = optimizer_factory(cfg["training"], model.parameters())
    for epoch in (0, 20):
        nio.save_checkpoints(epoch, model, opt, str(tmp_path))
    nio.save_best_checkpoints(7, model, str(tmp_path), 0.123456)
    names = sorted(os.listdir(tmp_path))
    assert names == ["model_00000", "model_00020", "modelbest_00007_0.123456", "opt_00000", "opt_00020"]
    # a raw state_dict, as the reference writes it (and as build_model(weight_file=...) reads it)
    blob = torch.load(tmp_path / "model_00020")
    assert list(blob.keys()) == list(model.state_dict().keys())
    model2, *_ = build_model(cfg, weight_file=str(tmp_path / "model_00020"))
    with torch.no_grad():
        for p in model2.parameters():
            p.add_(1.0)
    _, opt2 = optimizer_factory(cfg["training"], model2.parameters())
    args = types.SimpleNamespace(continue_from_epoch=0)
    nio.load_checkpoints(model2, opt2, str(tmp_path), args, "cpu")
    assert args.continue_from_epoch == 21
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
    args = types.SimpleNamespace(continue_from_epoch=0, best_val_loss=None)
    nio.load_best_checkpoints(model2, str(tmp_path), args, "cpu")
    assert args.continue_from_epoch == 8 and abs(args.best_val_loss - 0.123456) < 1e-9


def test_resume_is_a_no_op_without_a_matching_optimizer_file(tmp_path):
    cfg = synth.make_config("backward")
    model, *_ = build_model(cfg)
    torch.save(model.state_dict(), tmp_path / "model_00003")
    _, opt = optimizer_factory(cfg["training"], model.parameters())
    args = types.SimpleNamespace(continue_from_epoch=0)
    nio.load_checkpoints(model, opt, str(tmp_path), args, "cpu")
    assert args.continue_from_epoch == 0


def test_stager_refuses_cpu():
    with pytest.raises(RuntimeError, match="CPU not supported"):
        nio.DeviceStager([], "cpu")


@pytest.mark.gpu
def test_stager_yields_every_batch_in_order_on_the_device():
    g = torch.Generator().manual_seed(0)
    batches = [{"space_samples_src": torch.rand(2, 500, 3, generator=g),
                "surface_samples_inputs": torch.rand(2, 256, 7, generator=g).half(),
                "idx": torch.tensor([i])} for i in range(7)]
    st = nio.DeviceStager(batches, "cuda:0")
    seen = 0
    for i, dev in enumerate(st):
        assert dev["space_samples_src"].is_cuda and dev["surface_samples_inputs"].dtype == torch.float32
        # consume on the current stream, like a training step would
        assert torch.equal(dev["space_samples_src"].cpu(), batches[i]["space_samples_src"])
        assert torch.equal(dev["surface_samples_inputs"].cpu(), batches[i]["surface_samples_inputs"].float())
        assert int(dev["idx"].item()) == i
        seen += 1
    assert seen == 7 and len(st._pinned[0]) == 3
    assert st.h2d_bytes == 7 * (2 * 500 * 3 * 4 + 2 * 256 * 7 * 2 + 8)


@pytest.mark.gpu
def test_stager_feeds_a_training_step():
    cfg = synth.make_config("forward")
    model, train_on_batch, _, _ = build_model(cfg, device="cuda:0")
    schema = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(synth.named_state_dict(schema, seed=0))
    model.train()
    _, opt = optimizer_factory(cfg["training"], model.parameters())
    batches = [synth.forward_batch(2, 1024, 2048, seed=s) for s in (1, 2, 3)]
    losses = [train_on_batch(model, opt, dev, cfg) for dev in nio.DeviceStager(batches, "cuda:0")]
    assert len(losses) == 3 and all(np.isfinite(l) for l in losses)
