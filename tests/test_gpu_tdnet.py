"""End-to-end GPU parity of the product model (nsdp_b200.model, CUDA kernels through the C ABI) against
(a) the golden vectors produced by the live reference (tests/golden/) and (b) the CPU oracle."""
import numpy as np
import pytest
import torch

from nsdp_b200 import synth
from nsdp_b200.model import build_model
from oracle import tdnet_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4  # BASELINE.json north_star: flow L2 error vs reference < 1e-4 (mean per-point L2)


def _mean_l2(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64), axis=-1).mean())


def _model(schemas, mtype):
    model, *_ = build_model(synth.make_config(mtype), device=DEV)
    model.load_state_dict(synth.named_state_dict([(k, s) for k, s in schemas[mtype]], seed=0))
    return model


@pytest.mark.parametrize("mtype", ["forward", "backward"])
def test_c1_forward_against_reference_golden(golden, schemas, mtype):
    model = _model(schemas, mtype).eval()
    batch = synth.forward_batch(1, 1024, 2048, seed=1234, fp16_grid=False)
    with torch.no_grad():
        surf = batch["surface_samples_inputs"].to(DEV)
        enc = model.encode(surf)
        out = model(batch["space_samples_src"].to(DEV), surf)
    np.testing.assert_array_equal(enc["anchors"].cpu().numpy(), golden[f"c1_{mtype}_anchors"])  # FPS: bit-exact
    np.testing.assert_allclose(enc["z"].cpu().numpy(), golden[f"c1_{mtype}_z"], atol=1e-4, rtol=1e-3)
    err = _mean_l2(out.cpu().numpy(), golden[f"c1_{mtype}_flow"])
    assert err < TOL, err


def test_flow_arbitrary_eval_against_reference_golden(golden, schemas):
    model = _model(schemas, "arbitrary").eval()
    b = synth.forward_batch(1, 1024, 512, seed=77, fp16_grid=False)
    s = b["surface_samples_inputs"].to(DEV)
    src = s[:, :, 0:3].contiguous()
    with torch.no_grad():
        space_c = model.model_canonicalize(b["space_samples_src"].to(DEV), src)
        surf_c = model.model_canonicalize(src, src)
        assert _mean_l2(space_c.cpu().numpy(), golden["arb_eval_space_src2cano"]) < TOL
        assert _mean_l2(surf_c.cpu().numpy(), golden["arb_eval_surface_src2cano"]) < TOL
        # stage 2 teacher-forced with the reference's stage-1 coordinates (see tests/test_oracle_golden.py)
        inp = torch.cat([torch.from_numpy(golden["arb_eval_surface_src2cano"]).to(DEV), s[:, :, 3:6], s[:, :, 6:7]], -1)
        out_tf = model.model_deform(torch.from_numpy(golden["arb_eval_space_src2cano"]).to(DEV), inp.contiguous())
        assert _mean_l2(out_tf.cpu().numpy(), golden["arb_eval_flow"]) < TOL
        out = model(b["space_samples_src"].to(DEV), src, s[:, :, 3:6], s[:, :, 6:7])
        assert _mean_l2(out.cpu().numpy(), golden["arb_eval_flow"]) < 5e-3


def test_forward_against_oracle_multi_shape_fp16_grid(schemas):
    """B=3 shapes on the fp16 grid (real data layout): anchors bit-exact, flow within tolerance of the CPU oracle."""
    model = _model(schemas, "forward").eval()
    batch = synth.forward_batch(3, 1500, 1000, seed=42, fp16_grid=True)
    cfg = synth.make_config("forward")["model"]
    sd = synth.named_state_dict([(k, s) for k, s in schemas["forward"]], seed=0)
    trace = {}
    with torch.no_grad():
        want = orc.tdnet_forward(sd, "", batch["space_samples_src"], batch["surface_samples_inputs"], cfg, False, trace=trace)
        surf = batch["surface_samples_inputs"].to(DEV)
        enc = model.encode(surf)
        got = model.decode(batch["space_samples_src"].to(DEV), enc)
    assert torch.equal(enc["anchors"].cpu(), trace["anchors"])
    assert _mean_l2(got.cpu().numpy(), want.numpy()) < TOL


def test_forward_full_size_shape_against_oracle(schemas):
    """One shape at BASELINE.json's full size (4096 surface points on the fp16 grid, 50 000 spatial queries): anchors
    bit-exact (two FPS levels on 4096 points with real ties), flow within the 1e-4 tolerance of the CPU oracle. The batch
    of 8 such shapes of the bench differs only by the batch dimension (BatchNorm is in eval mode here)."""
    model = _model(schemas, "forward").eval()
    batch = synth.forward_batch(1, 4096, 50000, seed=1234, fp16_grid=True)
    cfg = synth.make_config("forward")["model"]
    sd = synth.named_state_dict([(k, s) for k, s in schemas["forward"]], seed=0)
    trace = {}
    with torch.no_grad():
        want = orc.tdnet_forward(sd, "", batch["space_samples_src"], batch["surface_samples_inputs"], cfg, False, trace=trace)
        surf = batch["surface_samples_inputs"].to(DEV)
        enc = model.encode(surf)
        got = model.decode(batch["space_samples_src"].to(DEV), enc)
    assert torch.equal(enc["anchors"].cpu(), trace["anchors"])
    assert _mean_l2(got.cpu().numpy(), want.numpy()) < TOL


@pytest.mark.parametrize("impl", [1, 0], ids=["fp32-cuda-cores", "auto-tcgen05"])
def test_training_step_against_reference_golden(golden, golden_r2, schemas, impl, monkeypatch):
    """fwd + bwd through the CUDA kernels (train-mode BatchNorm): loss, prediction, d/d query coordinates,
    d/d surface inputs, selected parameter gradients, all gradient norms and BN running stats vs the live fp32
    reference (tests/golden/make_golden.py, 'train_fwd_*').

    Gradient bar: relative L2 < 1e-3 on BOTH kernel families. d/d query is compared on the queries that do
    not sit on a ReLU kink (`fw64_keep`, the 70-odd % whose smallest decoder pre-activation is > 1e-4 of the layer rms):
    a gradient is discontinuous there, and two fp32 implementations of the reference itself differ by 2e-3 on the full
    tensor and by 1e-6 on the kept rows (tests/golden/make_golden_r2.py). The kink rows are excluded on both sides."""
    from nsdp_b200 import ops
    monkeypatch.setattr(ops, "VATTN_IMPL", impl)
    monkeypatch.setattr(ops, "TAIL_IMPL", impl)
    gtol = 1e-3
    keep = golden_r2["fw64_keep"]
    model = _model(schemas, "forward").train()
    b = synth.forward_batch(2, 768, 640, seed=5, fp16_grid=False)
    q = b["space_samples_src"].to(DEV).requires_grad_(True)
    surf = b["surface_samples_inputs"].to(DEV).requires_grad_(True)
    pred = model(q, surf)
    from nsdp_b200.model.utils import compute_l2_error
    loss = compute_l2_error(pred, b["space_samples_tgt"].to(DEV))
    loss.backward()
    assert abs(loss.item() - float(golden["train_fwd_loss"])) < 1e-5
    assert _mean_l2(pred.detach().cpu().numpy(), golden["train_fwd_pred"]) < TOL
    rel = lambda a, ref: float(np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30))
    assert rel(q.grad.cpu().numpy()[keep], golden["train_fwd_dq"][keep]) < gtol
    assert rel(surf.grad.cpu().numpy(), golden["train_fwd_dsurf"]) < gtol
    grads = {k: p.grad for k, p in model.named_parameters()}
    for key in golden.files:
        if key.startswith("train_fwd_grad::"):
            k = key.split("::", 1)[1]
            assert rel(grads[k].cpu().numpy(), golden[key]) < gtol, k
        if key.startswith("train_fwd_buf::"):
            k = key.split("::", 1)[1]
            np.testing.assert_allclose(model.state_dict()[k].cpu().numpy(), golden[key], atol=1e-5, rtol=1e-4)
    norms = np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in model.named_parameters()])
    ref = golden["train_fwd_gradnorms"]
    names = [k for k, _ in model.named_parameters()]
    for n, a, r in zip(names, norms, ref):
        if n.endswith("fc_gamma.2.bias") or n.endswith("fc_gamma1.2.bias") or n.endswith("fc_gamma2.2.bias"):
            # constant over the softmax axis: the true gradient is exactly 0; the reference's value is rounding
            # noise (autograd through softmax), ours is exactly 0 / None
            assert r < 1e-6 and (a <= 1e-6)
            continue
        # parameters feeding straight into a BatchNorm (e.g. a bias) have a true gradient of 0: both sides are noise
        assert abs(a - r) <= 2 * gtol * r + 1e-7, (n, a, r)


@pytest.mark.parametrize("impl", [1, 0], ids=["fp32-cuda-cores", "auto-tcgen05"])
def test_training_step_every_gradient_against_fp64_truth(golden_r2, schemas, impl, monkeypatch):
    """The same step against the fp64 run of the live reference, kink queries masked out of the LOSS on both sides
    (make_golden_r2.py part 2): d/d query, d/d surface and EVERY parameter gradient within 1e-3 relative L2 (tensors kept
    in full) / 1.5e-3 (seeded-projection estimate), or 3x the reference's own fp32 distance from the truth where that is
    larger (ill-conditioned BatchNorm backward)."""
    from helpers_r2 import check_param_grads, masked_l2, rel
    from nsdp_b200 import ops
    monkeypatch.setattr(ops, "VATTN_IMPL", impl)
    monkeypatch.setattr(ops, "TAIL_IMPL", impl)
    g = golden_r2
    keep = g["fw64_keep"]
    model = _model(schemas, "forward").train()
    b = synth.forward_batch(2, 768, 640, seed=5, fp16_grid=False)
    q = b["space_samples_src"].to(DEV).requires_grad_(True)
    surf = b["surface_samples_inputs"].to(DEV).requires_grad_(True)
    pred = model(q, surf)
    loss = masked_l2(pred, b["space_samples_tgt"].to(DEV), torch.from_numpy(keep).to(DEV))
    loss.backward()
    assert abs(loss.item() - float(g["fw64_loss"])) < 1e-5
    assert _mean_l2(pred.detach().cpu().numpy(), g["fw64_pred"]) < TOL
    e_dq = rel(q.grad.cpu().numpy()[keep], g["fw64_dq"][keep])
    e_ds = rel(surf.grad.cpu().numpy(), g["fw64_dsurf"])
    assert e_dq < 1e-3 and e_ds < 1e-3, (e_dq, e_ds)
    grads = {k: (None if p.grad is None else p.grad.cpu().numpy()) for k, p in model.named_parameters()}
    # tensors kept in full: 1e-3 on both kernel families. Projection-estimated tensors (16 projections: the estimate itself
    # scatters by ~18 %): 3e-3. ReLU mask flips inside the ENCODER (2.5 M pre-activations per shape in a full-attention block)
    # cannot be left out of the loss, and ANY change of fp32 rounding re-rolls them: on encoder.final_transformers.0.fc_gamma.0
    # the reference's own fp32 run is 3e-4 from the truth, the fp32 CUDA-core kernels 8e-4 .. 1.7e-3 depending on how the first
    # block's projections are associated, the tcgen05 kernels (bf16x3: ~3e-6 on every pre-activation) 1.7e-3.
    # tools/kink_noise_emulation.py reproduces the effect on the CPU: the fp32 oracle plus 3e-6 noise on the pre-activations
    # lands at 0.6 - 1.3e-3 on the same tensors (measured).
    worst = check_param_grads(grads, g, "fw64", 1e-3, 3e-3)
    print(f"[impl {impl}] d/dq {e_dq:.2e} (reference fp32: {float(g['fw64_ref32err_dq']):.2e}), d/dsurf {e_ds:.2e} "
          f"(reference fp32: {float(g['fw64_ref32err_dsurf']):.2e}), worst parameter gradient {worst}")


def test_flow_arbitrary_staged_training_step_against_fp64_truth(golden_r2, schemas):
    """FlowArbitrary fwd+bwd stage by stage (make_golden_r2.py part 3): (a) stage 1 free-running vs the fp32 reference;
    (b) stage 2 teacher-forced with the reference's stage-1 outputs vs fp64 truth: loss, prediction, the gradients that
    reach the stage-1 outputs, every deform-net parameter gradient; (c) stage-1 backward driven by the truth's upstream
    gradients: every canonicalise-net parameter gradient. Bars as in the test above."""
    from helpers_r2 import check_param_grads, masked_l2, rel
    g = golden_r2
    model = _model(schemas, "arbitrary").train()
    b = synth.forward_batch(2, 640, 384, seed=9, fp16_grid=False)
    s = b["surface_samples_inputs"].to(DEV)
    src, tgt, mask = s[:, :, 0:3].contiguous(), s[:, :, 3:6], s[:, :, 6:7]
    cano, deform = model.model_canonicalize, model.model_deform
    enc = cano.encode(src)
    space_c = cano.decode(b["space_samples_src"].to(DEV), enc)
    surf_c = cano.decode(src, enc)
    assert _mean_l2(space_c.detach().cpu().numpy(), g["arb_space_src2cano"]) < TOL
    assert _mean_l2(surf_c.detach().cpu().numpy(), g["arb_surface_src2cano"]) < TOL
    space_in = torch.from_numpy(g["arb_space_src2cano"]).to(DEV).requires_grad_(True)
    surf_in = torch.from_numpy(g["arb_surface_src2cano"]).to(DEV).requires_grad_(True)
    pred = deform(space_in, torch.cat([surf_in, tgt, mask], -1).contiguous())
    assert _mean_l2(pred.detach().cpu().numpy(), g["arb_pred"]) < TOL
    keep2 = g["arb_keep2"]
    loss = masked_l2(pred, b["space_samples_tgt"].to(DEV), torch.from_numpy(keep2).to(DEV))
    assert abs(loss.item() - float(g["arb_loss"])) < 1e-5
    loss.backward()
    e_sp = rel(space_in.grad.cpu().numpy()[keep2], g["arb_d_space_src2cano"][keep2])
    e_su = rel(surf_in.grad.cpu().numpy(), g["arb_d_surface_src2cano"])
    assert e_sp < 1e-3 and e_su < max(1e-3, 3 * float(g["arb_ref32err_d_surface"])), (e_sp, e_su)
    grads = {k: (None if p.grad is None else p.grad.cpu().numpy()) for k, p in model.named_parameters()
             if k.startswith("model_deform.")}
    # tcgen05 kernels: see the test above for the kink-flip floor of the projection-estimated encoder tensors; here the
    # gradient additionally enters the encoder through the decoder's fp16-staged table gradients (2e-4) and train-mode
    # BatchNorm backward (cancellation; the reference's own fp32 is at 1e-2 on some of these tensors): measured worst
    # 2.3e-3 (bf16x2 staging) / 3.0e-3 (fp16 staging), bar 5e-3; tensors kept in full stay at 1e-3
    w2 = check_param_grads(grads, g, "arb2", 1e-3, 5e-3)
    up_s = torch.from_numpy(g["arb_d_space_src2cano"] * g["arb_keep1_space"][..., None]).to(DEV)
    up_f = torch.from_numpy(g["arb_d_surface_src2cano"] * g["arb_keep1_surface"][..., None]).to(DEV)
    torch.autograd.backward([space_c, surf_c], [up_s, up_f])
    grads = {k: (None if p.grad is None else p.grad.cpu().numpy()) for k, p in model.named_parameters()
             if k.startswith("model_canonicalize.")}
    w1 = check_param_grads(grads, g, "arb1", 1e-3, 5e-3)
    print(f"stage 2: d/d space {e_sp:.2e}, d/d surface {e_su:.2e}, worst deform gradient {w2}; stage 1: worst {w1}")


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_encoder_per_block_activations_against_reference(golden_r2, schemas, mode):
    """the output of every encoder block (transformer_begin, each TransitionDown / ElementwiseMLP /
    TransformerBlock, the three final blocks) against forward-hook captures of the live reference
    (make_golden_r2.py part 1) — eval mode on the C1 cloud, train mode (batch-statistics BatchNorm) on the training batch."""
    from helpers_r2 import rel, thin
    model = _model(schemas, "forward")
    model = model.eval() if mode == "eval" else model.train()
    b = synth.forward_batch(1, 1024, 2048, seed=1234, fp16_grid=False) if mode == "eval" else \
        synth.forward_batch(2, 768, 640, seed=5, fp16_grid=False)
    got, hooks = {}, []
    for name, mod in model.encoder.named_modules():
        if f"trace_{mode}::{name}" in golden_r2.files:
            def hook(_m, _i, o, name=name):
                got[name] = thin((o[1] if isinstance(o, tuple) else o).detach().cpu().numpy())
            hooks.append(mod.register_forward_hook(hook))
    with torch.no_grad():
        model.encoder(b["surface_samples_inputs"].to(DEV))
    for h in hooks:
        h.remove()
    keys = [k.split("::", 1)[1] for k in golden_r2.files if k.startswith(f"trace_{mode}::")]
    assert len(keys) == 15 and set(keys) == set(got)
    for k in keys:
        e = rel(got[k], golden_r2[f"trace_{mode}::{k}"])
        assert e < 1e-4, (k, e)


def test_flow_arbitrary_training_step(golden, schemas):
    """FlowArbitrary fwd+bwd: gradient flows through query AND surface coordinates into the canonicalise net.
    Compared loosely with the reference (stage-2 FPS/k-NN decisions are discontinuous in stage-1 outputs);
    BatchNorm bookkeeping of the encode-once trick is compared exactly."""
    model = _model(schemas, "arbitrary").train()
    b = synth.forward_batch(2, 640, 384, seed=9, fp16_grid=False)
    s = b["surface_samples_inputs"].to(DEV)
    pred = model(b["space_samples_src"].to(DEV), s[:, :, 0:3], s[:, :, 3:6], s[:, :, 6:7])
    from nsdp_b200.model.utils import compute_l2_error
    loss = compute_l2_error(pred, b["space_samples_tgt"].to(DEV))
    loss.backward()
    assert abs(loss.item() - float(golden["train_arb_loss"])) < 2e-2 * float(golden["train_arb_loss"])
    sd = model.state_dict()
    k = "model_canonicalize.encoder.transformer_begin.bn."
    assert int(sd[k + "num_batches_tracked"]) == int(golden["train_arb_buf::" + k + "num_batches_tracked"]) == 2
    np.testing.assert_allclose(sd[k + "running_mean"].cpu().numpy(), golden["train_arb_buf::" + k + "running_mean"],
                               atol=1e-5, rtol=1e-4)
    norms = np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in model.named_parameters()])
    ref = golden["train_arb_gradnorms"]
    has = ref > 1e-6
    # every parameter the reference trains receives a gradient here too, of comparable size
    assert np.all(norms[has] > 0)
    ratio = norms[has] / ref[has]
    assert np.median(np.abs(ratio - 1)) < 2e-2, np.median(np.abs(ratio - 1))
    # canonicalise-net parameters get their gradient only through coordinates (d/d xyz_q and d/d surface xyz)
    names = [n for n, _ in model.named_parameters()]
    idx = names.index("model_canonicalize.decoder.fc_out.weight")
    assert norms[idx] > 0 and abs(norms[idx] / ref[idx] - 1) < 5e-2


def test_flow_arbitrary_full_size_step_and_stagewise_forward(schemas):
    """BASELINE configs[2] at its real per-GPU size: FlowArbitrary with 4 shapes x 4096 surface points x 50 000 queries.
    (a) one training step runs (finite loss, every trained parameter receives a finite gradient, peak memory far below the
    reference's ~40 GB of saved activations); (b) shape 0, eval mode, against the CPU oracle stage by stage: canonicalised
    queries / surface free-running, deformation teacher-forced with the oracle's stage-1 coordinates (FPS / k-NN on stage-1
    OUTPUTS are discontinuous), all within the 1e-4 flow tolerance."""
    from nsdp_b200.model.utils import compute_l2_error
    B, N, Q = 4, 4096, 50000
    model = _model(schemas, "arbitrary").train()
    b = synth.forward_batch(B, N, Q, seed=1234, fp16_grid=True)
    s = b["surface_samples_inputs"].to(DEV)
    src, tgt, mask = s[:, :, 0:3], s[:, :, 3:6], s[:, :, 6:7]
    torch.cuda.reset_peak_memory_stats()
    pred = model(b["space_samples_src"].to(DEV), src, tgt, mask)
    assert pred.shape == (B, Q, 3)
    loss = compute_l2_error(pred, b["space_samples_tgt"].to(DEV))
    loss.backward()
    assert np.isfinite(loss.item())
    # without a gradient: only the canonicalise net's pos_only first block's unused q/k/v weights (model/encoder/blocks.py:
    # 119,126) and the fc_gamma*.2 biases, which cancel in the softmax (exactly 0 here, rounding noise in the reference)
    missing = [k for k, p in model.named_parameters() if p.grad is None and not (
        any(t in k for t in ("model_canonicalize.encoder.transformer_begin.w_qs", "model_canonicalize.encoder.transformer_begin.w_ks",
                             "model_canonicalize.encoder.transformer_begin.w_vs")) or k.endswith(".2.bias"))]
    assert not missing, missing[:5]
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    assert torch.cuda.max_memory_allocated() < 20 * 2 ** 30
    del pred, loss
    model.zero_grad(set_to_none=True)
    # ---- (b) stage-wise forward of shape 0 against the oracle ------------------------------------------------------------
    cfg = synth.make_config("arbitrary")["model"]
    sd = synth.named_state_dict([(k, s_) for k, s_ in schemas["arbitrary"]], seed=0)
    model.load_state_dict(sd)          # the training-mode pass above moved the BatchNorm running statistics
    model.eval()
    q0, src0 = b["space_samples_src"][:1], b["surface_samples_inputs"][:1, :, 0:3].contiguous()
    rest0 = b["surface_samples_inputs"][:1, :, 3:7]
    with torch.no_grad():
        want_space = orc.tdnet_forward(sd, "model_canonicalize.", q0, src0, cfg, True)
        want_surf = orc.tdnet_forward(sd, "model_canonicalize.", src0, src0, cfg, True)
        want_flow = orc.tdnet_forward(sd, "model_deform.", want_space, torch.cat([want_surf, rest0], -1).contiguous(), cfg, False)
        cano, deform = model.model_canonicalize, model.model_deform
        enc = cano.encode(src0.to(DEV))
        got_space = cano.decode(q0.to(DEV), enc)
        got_surf = cano.decode(src0.to(DEV), enc)
        got_flow = deform(want_space.to(DEV), torch.cat([want_surf, rest0], -1).contiguous().to(DEV))
    assert _mean_l2(got_space.cpu().numpy(), want_space.numpy()) < TOL
    assert _mean_l2(got_surf.cpu().numpy(), want_surf.numpy()) < TOL
    assert _mean_l2(got_flow.cpu().numpy(), want_flow.numpy()) < TOL


def test_non_default_configuration_forward_and_gradients(golden_r2):
    """Every config knob the reference's blocks honour in ONE non-default model (synth.make_alt_config: three
    down-sampling levels, local final attention with k = 2 * nneighbor, widths 64 / 128 / 96, 5 neighbours, 3 ResNet blocks):
    the kernels' shape dispatch away from the bench configuration. Forward vs the live-reference fixture (anchors bit-exact,
    flow 1e-4); training step vs the oracle's autograd on the CPU (loss, d/d query on non-kink rows, selected gradients)."""
    from helpers_r2 import masked_l2, rel
    g = golden_r2
    schema = [(str(k), tuple(int(x) for x in str(s).split(",") if x)) for k, s in zip(g["alt_schema_keys"], g["alt_schema_shapes"])]
    cfg = synth.make_alt_config()
    model, *_ = build_model(cfg, device=DEV)
    assert [(k, tuple(v.shape)) for k, v in model.state_dict().items()] == schema        # same schema as the reference builds
    sd = synth.named_state_dict(schema, seed=4)
    model.load_state_dict(sd)
    model.eval()
    b = synth.forward_batch(2, 1500, 300, seed=31, fp16_grid=True)
    with torch.no_grad():
        surf = b["surface_samples_inputs"].to(DEV)
        enc = model.encode(surf)
        out = model.decode(b["space_samples_src"].to(DEV), enc)
    np.testing.assert_array_equal(enc["anchors"].cpu().numpy(), g["alt_anchors"])
    np.testing.assert_allclose(enc["z"].cpu().numpy(), g["alt_z"], atol=1e-4, rtol=1e-3)
    assert _mean_l2(out.cpu().numpy(), g["alt_flow"]) < TOL
    # training step: fp64 oracle as the truth, kink queries masked out of the loss on both sides
    sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    params = {k: v.requires_grad_(True) for k, v in sd64.items() if k.rsplit(".", 1)[-1] in ("weight", "bias")}
    kink = {}
    with torch.no_grad():
        orc.tdnet_forward({k: v.detach().clone() for k, v in sd64.items()}, "", b["space_samples_src"].double(),
                          b["surface_samples_inputs"].double(), cfg["model"], False, training=True, kink=kink)
    keep = kink["margin"] > 1e-4
    q64 = b["space_samples_src"].double().requires_grad_(True)
    want = orc.tdnet_forward(sd64, "", q64, b["surface_samples_inputs"].double(), cfg["model"], False, training=True)
    loss64 = masked_l2(want, b["space_samples_tgt"].double(), keep)
    loss64.backward()
    model.train()
    q = b["space_samples_src"].to(DEV).requires_grad_(True)
    pred = model(q, surf)
    loss = masked_l2(pred, b["space_samples_tgt"].to(DEV), keep.to(DEV))
    loss.backward()
    assert abs(loss.item() - loss64.item()) < 1e-5
    assert _mean_l2(pred.detach().cpu().numpy(), want.detach().numpy()) < TOL
    k = keep.numpy()
    assert rel(q.grad.cpu().numpy()[k], q64.grad.numpy()[k]) < 1e-3
    grads = dict(model.named_parameters())
    for name in ("decoder.fc_out.weight", "decoder.ct1.fc_gamma.0.weight", "decoder.ct1.w_ks.weight", "decoder.blocks.2.fc_0.weight",
                 "encoder.fc_middle.0.weight", "encoder.final_transformers.1.w_vs.weight", "encoder.transition_downs.2.sa.fc_gamma2.0.weight",
                 "encoder.transformer_begin.fc_delta.2.weight", "encoder.enc_sdf.weight", "encoder.elementwise.2.conv1.weight"):
        e = rel(grads[name].grad.cpu().numpy(), params[name].grad.numpy())
        assert e < 3e-3, (name, e)
