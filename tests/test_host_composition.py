"""Host logic of the model mirror on the CPU: does `nsdp_b200/model/*` — the weight folds (fold_sites / fold_pair_mlps), the
fused per-point projections, the sign conventions, the global-token algebra of the decoder, the packed tail weights, the
BatchNorm placement — compose the KERNEL CONTRACTS of include/nsdp_b200.h into the reference's mathematics?

The `-m gpu` kernel tests (tests/test_gpu_vattn.py, test_gpu_emlp.py, ...) hold every CUDA kernel to a float64 restatement of
its contract in the header. This file closes the other half without a GPU: the entry points of `nsdp_b200.ops` that the
model calls are replaced, FOR THIS TEST ONLY, by plain differentiable torch restatements of those same header contracts, and
the unmodified mirror modules then run on CPU tensors against (a) the vectors the LIVE reference produced
(tests/golden/tdnet_reference.npz, make_golden.py) and (b) the oracle in float64 (1e-9: pure algebra, no rounding slack).
kernel == contract (GPU tests) and contract + host code == reference (here) give kernel path == reference at block level.

Test infrastructure: the stand-ins live in tests/helpers/contracts.py, never in the package — the product has no CPU path (tests/test_abi.py)."""
import numpy as np
import pytest
import torch

from nsdp_b200 import synth
from nsdp_b200.model import build_model
from oracle import tdnet_oracle as orc


@pytest.fixture
def contracts(monkeypatch):
    """The header's kernel contracts (tests/helpers/contracts.py) stand in for the ops entry points, for this test only."""
    from helpers import contracts as hc
    hc.install(monkeypatch.setattr)


def _model(schemas, mtype, cfg=None, dtype=torch.float32):
    cfg = cfg or synth.make_config(mtype)
    model, *_ = build_model(cfg, device="cpu")
    schema = schemas[mtype] if schemas is not None else [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    sd = synth.named_state_dict([(k, s) for k, s in schema], seed=0)
    model.load_state_dict(sd)
    if dtype == torch.float64:
        model.double()
        sd = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    return model, sd, cfg


def _mean_l2(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64), axis=-1).mean())


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mtype", ["forward", "backward"])
def test_mirror_on_contracts_reproduces_the_live_reference_c1(contracts, golden, schemas, mtype):
    """BASELINE configs[0] (1 x 1024 surface points x 2048 queries): anchors bit-exact, flow within the 1e-4 bar of
    north_star by two orders of magnitude (fp32, different association of the folded weights only)."""
    model, _, _ = _model(schemas, mtype)
    model.eval()
    batch = synth.forward_batch(1, 1024, 2048, seed=1234, fp16_grid=False)
    with torch.no_grad():
        enc = model.encode(batch["surface_samples_inputs"])
        out = model.decode(batch["space_samples_src"], enc)
    np.testing.assert_array_equal(enc["anchors"].numpy(), golden[f"c1_{mtype}_anchors"])
    assert _mean_l2(out.numpy(), golden[f"c1_{mtype}_flow"]) < 2e-6
    np.testing.assert_allclose(enc["z"].numpy(), golden[f"c1_{mtype}_z"], atol=5e-5, rtol=1e-4)
    np.testing.assert_allclose(enc["anchor_feats"].numpy(), golden[f"c1_{mtype}_anchor_feats"], atol=1e-4, rtol=1e-4)


def test_mirror_on_contracts_equals_the_oracle_in_float64(contracts, schemas):
    """Pure algebra: in float64 the folded / fused host code and the reference's op-by-op chain agree to 1e-9, in eval AND in
    train mode (batch statistics), outputs AND gradients (queries, surface, every parameter that receives one)."""
    model, sd, cfg = _model(schemas, "forward", dtype=torch.float64)
    b = synth.forward_batch(2, 512, 300, seed=5, fp16_grid=False)
    for training in (False, True):
        model.train(training)
        ref_sd = {k: v.clone().requires_grad_(v.is_floating_point() and k.rsplit(".", 1)[-1] in ("weight", "bias"))
                  for k, v in sd.items()}
        q1 = b["space_samples_src"].double().requires_grad_(True)
        s1 = b["surface_samples_inputs"].double().requires_grad_(True)
        q2, s2 = q1.detach().clone().requires_grad_(True), s1.detach().clone().requires_grad_(True)
        model.zero_grad(set_to_none=True)
        out = model.decode(q1, model.encode(s1))
        want = orc.tdnet_forward(ref_sd, "", q2, s2, cfg["model"], no_input_corr=False, training=training)
        assert float((out - want).detach().abs().max()) < 1e-9 * max(1.0, float(want.detach().abs().max()))
        tgt = b["space_samples_tgt"].double()
        orc.l2_loss(out, tgt).backward()
        orc.l2_loss(want, tgt).backward()
        rel = lambda a, r: float((a - r).norm() / r.norm().clamp_min(1e-300))
        assert rel(q1.grad, q2.grad) < 1e-9 and rel(s1.grad, s2.grad) < 1e-9
        checked = 0
        for name, p in model.named_parameters():
            g = ref_sd[name].grad
            if g is None or float(g.norm()) < 1e-12 or p.grad is None:
                # mathematically zero gradients: the unused q/k/v weights of nothing here, but fc_gamma[2].bias (cancels in the
                # softmax: the kernel contract does not even take it), biases in front of a train-mode BatchNorm
                assert g is None or float(g.norm()) < 1e-12, (name, training)
                assert p.grad is None or float(p.grad.norm()) < 1e-12, (name, training)
                continue
            assert rel(p.grad, g) < 1e-8, (name, training, rel(p.grad, g))
            checked += 1
        assert checked > 200, checked


def test_mirror_on_contracts_training_step_matches_the_live_reference(contracts, golden, schemas):
    """The fp32 training-step vectors of the live reference (train-mode BatchNorm, loss, d/d query, d/d surface, selected
    parameter gradients, running statistics): same bars as the oracle's own pin (tests/test_oracle_golden.py)."""
    model, _, _ = _model(schemas, "forward")
    model.train()
    b = synth.forward_batch(2, 768, 640, seed=5, fp16_grid=False)
    q = b["space_samples_src"].clone().requires_grad_(True)
    surf = b["surface_samples_inputs"].clone().requires_grad_(True)
    pred = model.decode(q, model.encode(surf))
    loss = orc.l2_loss(pred, b["space_samples_tgt"])
    loss.backward()
    assert abs(loss.item() - float(golden["train_fwd_loss"])) < 1e-6
    assert _mean_l2(pred.detach().numpy(), golden["train_fwd_pred"]) < 2e-6
    rel = lambda a, r: np.linalg.norm(a - r) / max(np.linalg.norm(r), 1e-30)
    assert rel(q.grad.numpy(), golden["train_fwd_dq"]) < 1e-3
    assert rel(surf.grad.numpy(), golden["train_fwd_dsurf"]) < 1e-3
    params = dict(model.named_parameters())
    state = model.state_dict()
    n_grad = n_buf = 0
    for key in golden.files:
        if key.startswith("train_fwd_grad::"):
            k = key.split("::", 1)[1]
            assert rel(params[k].grad.numpy(), golden[key]) < 1e-3, k
            n_grad += 1
        if key.startswith("train_fwd_buf::"):
            k = key.split("::", 1)[1]
            np.testing.assert_allclose(state[k].numpy(), golden[key], atol=1e-6, rtol=1e-5)
            n_buf += 1
    assert n_grad >= 10 and n_buf >= 4


def test_non_default_configuration_on_contracts(contracts, golden_r2):
    """The knobs no shipped YAML varies (three levels, local final attention, other widths, 5 neighbours, 3 ResNet blocks): the
    host code's shape handling against the live reference's vectors (`alt_*`, tests/golden/make_golden_r2.py)."""
    g = golden_r2
    schema = [(str(k), tuple(int(x) for x in str(s).split(",") if x)) for k, s in zip(g["alt_schema_keys"], g["alt_schema_shapes"])]
    model, *_ = build_model(synth.make_alt_config(), device="cpu")
    assert [(k, tuple(v.shape)) for k, v in model.state_dict().items()] == schema
    model.load_state_dict(synth.named_state_dict(schema, seed=4))
    model.eval()
    b = synth.forward_batch(2, 1500, 300, seed=31, fp16_grid=True)
    with torch.no_grad():
        enc = model.encode(b["surface_samples_inputs"])
        out = model.decode(b["space_samples_src"], enc)
    np.testing.assert_array_equal(enc["anchors"].numpy(), g["alt_anchors"])
    np.testing.assert_allclose(enc["z"].numpy(), g["alt_z"], atol=5e-5, rtol=1e-4)
    assert _mean_l2(out.numpy(), g["alt_flow"]) < 2e-6


# ---------------------------------------------------------------------------------------------------------------------
# FlowArbitrary (model/flow_arbitrary.py:7-85): the mirror's encode-once composition
# ---------------------------------------------------------------------------------------------------------------------
def test_flow_arbitrary_eval_on_contracts(contracts, golden, schemas):
    """Stage 1 tight, stage 2 teacher-forced with the reference's stage-1 coordinates tight, free-running loose (FPS / k-NN of
    stage 2 are discontinuous in stage-1 outputs) — the protocol of tests/test_oracle_golden.py, on the mirror's host code."""
    model, _, _ = _model(schemas, "arbitrary")
    model.eval()
    b = synth.forward_batch(1, 1024, 512, seed=77, fp16_grid=False)
    s = b["surface_samples_inputs"]
    src, tgt, mask = s[:, :, 0:3].contiguous(), s[:, :, 3:6], s[:, :, 6:7]
    with torch.no_grad():
        space_c, surf_c = model._canonicalize_twice(b["space_samples_src"], src)
        assert _mean_l2(space_c.numpy(), golden["arb_eval_space_src2cano"]) < 2e-6
        assert _mean_l2(surf_c.numpy(), golden["arb_eval_surface_src2cano"]) < 2e-6
        inp = torch.cat([torch.from_numpy(golden["arb_eval_surface_src2cano"]), tgt, mask], -1).contiguous()
        forced = model.model_deform(torch.from_numpy(golden["arb_eval_space_src2cano"]), inp)
        assert _mean_l2(forced.numpy(), golden["arb_eval_flow"]) < 4e-6
        free = model(b["space_samples_src"], src, tgt, mask)
        assert _mean_l2(free.numpy(), golden["arb_eval_flow"]) < 5e-3


def test_flow_arbitrary_inference_encodes_once_with_the_references_result(contracts, schemas):
    """test_on_batch_with_arbitrary runs 2 encoder passes where the reference runs 6 (flow_arbitrary.py:65-85): in eval mode the
    outputs must be those of the reference's two full model(...) calls, restated with the oracle. In float64, so that stage 2's
    FPS / k-NN decisions see the same stage-1 coordinates on both sides and the comparison is pure algebra."""
    from nsdp_b200.model.flow_arbitrary import test_on_batch_with_arbitrary as infer
    model, sd, cfg = _model(schemas, "arbitrary", dtype=torch.float64)
    model.eval()
    b = synth.forward_batch(1, 600, 200, seed=21, fp16_grid=False)
    s = b["surface_samples_inputs"].double()
    src, tgt, mask = s[:, :, 0:3].contiguous(), s[:, :, 3:6], s[:, :, 6:7]
    verts, verts_tgt = b["space_samples_src"].double(), b["space_samples_tgt"].double()
    calls = {"n": 0}
    for net in (model.model_canonicalize, model.model_deform):
        net.encoder.register_forward_hook(lambda *_: calls.__setitem__("n", calls["n"] + 1))
    data = {"surface_samples_inputs": s, "surface_samples_src": src, "verts_src": verts, "verts_tgt": verts_tgt}
    loss, out = infer(model, data, None, compute_loss=True)
    assert calls["n"] == 2
    with torch.no_grad():
        want_surf = orc.flow_arbitrary_forward(sd, src, src, tgt, mask, cfg["model"])
        want_verts = orc.flow_arbitrary_forward(sd, verts, src, tgt, mask, cfg["model"])
    assert float((out["surface_samples_tgt_pred"] - want_surf).abs().max()) < 1e-9
    assert float((out["verts_tgt_pred"] - want_verts).abs().max()) < 1e-9
    assert abs(loss - float(orc.l2_loss(want_verts, verts_tgt))) < 1e-12


def test_flow_arbitrary_training_bookkeeping_on_contracts(contracts, golden, schemas):
    """One canonicalise-encoder pass with momentum 2m - m^2 must leave the BatchNorm buffers the reference's TWO passes leave
    (flow_arbitrary.py:19-20), and the loss / prediction of the live reference's training forward."""
    model, _, _ = _model(schemas, "arbitrary")
    model.train()
    b = synth.forward_batch(2, 640, 384, seed=9, fp16_grid=False)
    s = b["surface_samples_inputs"]
    pred = model(b["space_samples_src"], s[:, :, 0:3], s[:, :, 3:6], s[:, :, 6:7])
    loss = orc.l2_loss(pred, b["space_samples_tgt"])
    loss.backward()
    assert abs(loss.item() - float(golden["train_arb_loss"])) < 2e-2 * float(golden["train_arb_loss"])
    state = model.state_dict()
    k = "model_canonicalize.encoder.transformer_begin.bn."
    assert int(state[k + "num_batches_tracked"]) == int(golden["train_arb_buf::" + k + "num_batches_tracked"]) == 2
    np.testing.assert_allclose(state[k + "running_mean"].numpy(), golden["train_arb_buf::" + k + "running_mean"],
                               atol=1e-6, rtol=1e-5)
    for bn in [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm1d)]:
        assert bn.momentum == 0.1                                   # restored after the pass
    norms = np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in model.named_parameters()])
    ref = golden["train_arb_gradnorms"]
    has = ref > 1e-6
    assert np.all(norms[has] > 0)                                   # every parameter the reference trains gets a gradient
    assert np.median(np.abs(norms[has] / ref[has] - 1)) < 2e-2


def test_forward_net_inference_encodes_once_with_the_references_result(contracts, schemas):
    """test_on_batch_with_cano (deformation_networks.py:90-109): one encoder pass for both query sets instead of the reference's
    two identical ones; same data_dict keys, same loss."""
    from nsdp_b200.model.deformation_networks import test_on_batch_with_cano as infer
    model, sd, cfg = _model(schemas, "forward", dtype=torch.float64)
    model.eval()
    b = synth.forward_batch(2, 400, 150, seed=23, fp16_grid=False)
    s = b["surface_samples_inputs"].double()
    verts, verts_tgt = b["space_samples_src"].double(), b["space_samples_tgt"].double()
    calls = {"n": 0}
    model.encoder.register_forward_hook(lambda *_: calls.__setitem__("n", calls["n"] + 1))
    data = {"surface_samples_inputs": s, "surface_samples_src": s[:, :, 0:3].contiguous(), "verts_src": verts, "verts_tgt": verts_tgt}
    loss, out = infer(model, data, None, compute_loss=True)
    assert calls["n"] == 1 and out is data
    with torch.no_grad():
        want_surf = orc.tdnet_forward(sd, "", data["surface_samples_src"], s, cfg["model"], False)
        want_verts = orc.tdnet_forward(sd, "", verts, s, cfg["model"], False)
    assert float((out["surface_samples_tgt_pred"] - want_surf).abs().max()) < 1e-9
    assert float((out["verts_tgt_pred"] - want_verts).abs().max()) < 1e-9
    assert abs(loss - float(orc.l2_loss(want_verts, verts_tgt))) < 1e-12
    loss0, _ = infer(model, data, None, compute_loss=False)
    assert loss0 == 0.0


# ---------------------------------------------------------------------------------------------------------------------
# ablation blocks: `encoder: pointnet++`, `decoder: interp`
# ---------------------------------------------------------------------------------------------------------------------
def test_ablation_model_on_contracts(contracts):
    import json
    import os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = np.load(os.path.join(gdir, "ablation_reference.npz"))
    with open(os.path.join(gdir, "ablation_schema.json")) as f:
        schema = json.load(f)
    model, *_ = build_model(synth.make_ablation_config(), device="cpu")
    model.load_state_dict(synth.named_state_dict([(k, s) for k, s in schema], seed=0))
    model.eval()
    batch = synth.forward_batch(1, 1024, 2048, seed=1234, fp16_grid=False)
    with torch.no_grad():
        enc = model.encode(batch["surface_samples_inputs"])
        out = model.decode(batch["space_samples_src"], enc)
    np.testing.assert_array_equal(enc["anchors"].numpy(), gold["c1_anchors"])
    np.testing.assert_allclose(enc["z"].numpy(), gold["c1_z"], atol=5e-5, rtol=1e-4)
    assert _mean_l2(out.numpy(), gold["c1_flow"]) < 2e-6
    model.train()
    b = synth.forward_batch(2, 768, 640, seed=5, fp16_grid=False)
    q = b["space_samples_src"].clone().requires_grad_(True)
    pred = model.decode(q, model.encode(b["surface_samples_inputs"]))
    loss = orc.l2_loss(pred, b["space_samples_tgt"])
    loss.backward()
    assert abs(loss.item() - float(gold["train_loss"])) < 1e-6
    assert _mean_l2(pred.detach().numpy(), gold["train_pred"]) < 2e-5      # fp32 through train-mode BatchNorm; GPU bar 1e-4
    # d/d query of an fp32 run vs another fp32 run: ReLU-kink flips dominate (no kink mask in this fixture)
    assert np.linalg.norm(q.grad.numpy() - gold["train_dq"]) / np.linalg.norm(gold["train_dq"]) < 1e-2
    norms = np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in model.named_parameters()])
    has = gold["train_gradnorms"] > 1e-6
    assert np.all(norms[has] > 0) and np.median(np.abs(norms[has] / gold["train_gradnorms"][has] - 1)) < 1e-3
