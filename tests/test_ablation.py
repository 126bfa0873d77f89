"""The reference's ablation blocks (`encoder: pointnet++`, `decoder: interp`) behind the same registries
(SURVEY.md section 8f, row 3): state_dict schema on the CPU, outputs and gradients on the GPU against fixtures minted
from the LIVE reference by tests/golden/make_golden_ablation.py."""
import json
import os

import numpy as np
import pytest
import torch

from nsdp_b200 import synth
from nsdp_b200.model import build_model
from nsdp_b200.model.decoder import decoder_dict
from nsdp_b200.model.encoder import encoder_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"


@pytest.fixture(scope="module")
def schema():
    with open(os.path.join(GOLDEN_DIR, "ablation_schema.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN_DIR, "ablation_reference.npz"))


def test_registries_match_the_reference():
    assert sorted(encoder_dict) == ["pointnet++", "pointransformer"]     # model/encoder/__init__.py:4-7
    assert sorted(decoder_dict) == ["crossatten", "interp"]              # model/decoder/__init__.py:5-8


def test_ablation_state_dict_schema_equals_reference(schema):
    model, *_ = build_model(synth.make_ablation_config())
    assert [[k, list(v.shape)] for k, v in model.state_dict().items()] == schema


def _mean_l2(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64), axis=-1).mean())


def _model(schema):
    model, *_ = build_model(synth.make_ablation_config(), device=DEV)
    model.load_state_dict(synth.named_state_dict([(k, s) for k, s in schema], seed=0))
    return model


@pytest.mark.gpu
def test_ablation_forward_against_reference_golden(schema, gold):
    model = _model(schema).eval()
    batch = synth.forward_batch(1, 1024, 2048, seed=1234, fp16_grid=False)
    with torch.no_grad():
        surf = batch["surface_samples_inputs"].to(DEV)
        enc = model.encode(surf)
        out = model(batch["space_samples_src"].to(DEV), surf)
    np.testing.assert_array_equal(enc["anchors"].cpu().numpy(), gold["c1_anchors"])          # FPS: bit-exact
    np.testing.assert_allclose(enc["z"].cpu().numpy(), gold["c1_z"], atol=1e-4, rtol=1e-3)
    assert _mean_l2(out.cpu().numpy(), gold["c1_flow"]) < 1e-4                                # BASELINE.json tolerance


@pytest.mark.gpu
def test_ablation_training_step_against_reference_golden(schema, gold):
    from nsdp_b200.model.utils import compute_l2_error
    model = _model(schema).train()
    b = synth.forward_batch(2, 768, 640, seed=5, fp16_grid=False)
    q = b["space_samples_src"].to(DEV).requires_grad_(True)
    pred = model(q, b["surface_samples_inputs"].to(DEV))
    loss = compute_l2_error(pred, b["space_samples_tgt"].to(DEV))
    loss.backward()
    assert abs(loss.item() - float(gold["train_loss"])) < 1e-5
    assert _mean_l2(pred.detach().cpu().numpy(), gold["train_pred"]) < 1e-4
    rel = lambda a, ref: float(np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30))
    assert rel(q.grad.cpu().numpy(), gold["train_dq"]) < 5e-3
    norms = np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in model.named_parameters()])
    ref = gold["train_gradnorms"]
    names = [k for k, _ in model.named_parameters()]
    for n, a, r in zip(names, norms, ref):
        if n.endswith("fc_gamma.2.bias"):   # constant over the softmax axis: true gradient 0 (reference: rounding noise)
            assert r < 1e-6 and a <= 1e-6
            continue
        if r < 0:   # no gradient in the reference either (the interpolation decoder never reads `z`: fc_middle is unused)
            assert a < 0 or a == 0.0, (n, a, r)
            continue
        # biases that feed straight into a BatchNorm have a true gradient of 0: both sides are rounding noise there, so the
        # absolute floor scales with the largest gradient of the model
        assert abs(a - r) <= 1e-2 * r + 1e-4 * float(ref.max()), (n, a, r)
