"""Generates the golden fixtures under tests/golden/ by running the LIVE reference (/root/reference/model,
unmodified) on the CPU in the authoring container. Run from the repo root:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed here as small .npz/.json files
together with this script. Inputs and weights are NOT stored: they are pure functions of (name, shape, seed)
from nsdp_b200/synth.py, so every consumer regenerates them bit-identically.

The only part of the reference that cannot execute on a CPU is its CUDA-only FPS kernel
(sampling.cpp:82-84); `pointnet2_ops._ext` is therefore shimmed with the C restatement in
oracle/nsdp_oracle.c (which in turn is pinned against the real kernel on the GPU box,
tests/test_gpu_index_kernels.py, and on the CPU against the reference's own torch FPS, tests/test_index_golden.py).
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from nsdp_b200 import synth  # noqa: E402
from oracle import tdnet_oracle as orc  # noqa: E402


def import_reference():
    ext = types.ModuleType("pointnet2_ops._ext")
    ext.furthest_point_sampling = lambda xyz, n: orc.fps(xyz, n)
    sys.modules["pointnet2_ops._ext"] = ext
    sys.path.insert(0, os.path.join(REF, "pointnet2_ops_lib"))
    sys.path.insert(0, REF)
    import model as ref_model  # noqa
    return ref_model


def schema_of(module):
    return [[k, list(v.shape)] for k, v in module.state_dict().items()]


def main():
    torch.set_num_threads(8)
    ref = import_reference()
    gold = {}
    schemas = {}

    # ---- state_dict schemas ------------------------------------------------------
    models = {}
    for mtype in ("forward", "backward", "arbitrary"):
        cfg = synth.make_config(mtype)
        m, *_ = ref.build_model(cfg)
        schemas[mtype] = schema_of(m)
        m.load_state_dict(synth.named_state_dict([(k, s) for k, s in schemas[mtype]], seed=0))
        models[mtype] = m
    with open(os.path.join(OUT, "state_dict_schema.json"), "w") as f:
        json.dump(schemas, f)

    # ---- C1: single-shape TDNet forward, eval mode (BASELINE.json configs[0]) ---------------------------
    batch = synth.forward_batch(1, 1024, 2048, seed=1234, fp16_grid=False)
    for mtype in ("forward", "backward"):
        m = models[mtype].eval()
        with torch.no_grad():
            enc_in = batch["surface_samples_inputs"]
            enc = m.encoder(enc_in[:, :, 0:3].contiguous()) if mtype == "backward" else m.encoder(enc_in)
            out = m(batch["space_samples_src"], enc_in)
        gold[f"c1_{mtype}_flow"] = out.numpy()
        gold[f"c1_{mtype}_z"] = enc["z"].numpy()
        gold[f"c1_{mtype}_anchors"] = enc["anchors"].numpy()
        gold[f"c1_{mtype}_anchor_feats"] = enc["anchor_feats"].numpy()

    # ---- FlowArbitrary forward, eval mode -----------------------------------------------------------------
    b2 = synth.forward_batch(1, 1024, 512, seed=77, fp16_grid=False)
    s = b2["surface_samples_inputs"]
    m = models["arbitrary"].eval()
    with torch.no_grad():
        out = m(b2["space_samples_src"], s[:, :, 0:3], s[:, :, 3:6], s[:, :, 6:7])
        # stage-1 outputs: FPS/k-NN inside stage 2 are discontinuous in these coordinates (a 1e-6 change flips
        # dozens of FPS picks), so consumers teacher-force stage 2 with the reference's own stage-1 result
        gold["arb_eval_space_src2cano"] = m.model_canonicalize(b2["space_samples_src"], s[:, :, 0:3]).numpy()
        gold["arb_eval_surface_src2cano"] = m.model_canonicalize(s[:, :, 0:3], s[:, :, 0:3]).numpy()
    gold["arb_eval_flow"] = out.numpy()

    # ---- training step (train-mode BN, loss, selected gradients, running stats), B=2 -------------------------
    b3 = synth.forward_batch(2, 768, 640, seed=5, fp16_grid=False)
    m = models["forward"].train()
    m.zero_grad()
    q = b3["space_samples_src"].clone().requires_grad_(True)
    surf = b3["surface_samples_inputs"].clone().requires_grad_(True)
    pred = m(q, surf)
    loss = ref.deformation_networks.compute_l2_error(pred, b3["space_samples_tgt"])
    loss.backward()
    gold["train_fwd_loss"] = np.array(loss.item(), np.float64)
    gold["train_fwd_pred"] = pred.detach().numpy()
    gold["train_fwd_dq"] = q.grad.numpy()
    gold["train_fwd_dsurf"] = surf.grad.numpy()
    grads = {k: p.grad for k, p in m.named_parameters()}
    for k in ("decoder.fc_out.weight", "decoder.blocks.2.fc_0.weight", "decoder.ct1.fc_gamma.0.weight",
              "decoder.ct1.fc_delta.0.weight", "decoder.ct1.w_ks.weight", "decoder.ct1.w_k_global.weight",
              "encoder.enc_sdf.weight", "encoder.transformer_begin.fc_delta.2.weight",
              "encoder.transition_downs.0.sa.fc_gamma2.0.weight", "encoder.final_transformers.1.w_vs.weight",
              "encoder.fc_middle.0.weight", "encoder.elementwise.1.conv1.weight"):
        gold["train_fwd_grad::" + k] = grads[k].numpy()
    gold["train_fwd_gradnorms"] = np.array([float(grads[k].norm()) if grads[k] is not None else -1.0
                                            for k, _ in m.named_parameters()], np.float64)
    sd = m.state_dict()
    for k in ("encoder.transformer_begin.bn.running_mean", "encoder.transformer_begin.bn.running_var",
              "encoder.final_elementwise.2.bn3.running_var", "encoder.transition_downs.1.sa.bnorm2.running_mean"):
        gold["train_fwd_buf::" + k] = sd[k].numpy().copy()

    # ---- FlowArbitrary training forward/backward (gradient flow through coordinates), B=1 ---------------------
    m = models["arbitrary"].train()
    m.zero_grad()
    b4 = synth.forward_batch(2, 640, 384, seed=9, fp16_grid=False)
    s = b4["surface_samples_inputs"]
    pred = m(b4["space_samples_src"], s[:, :, 0:3], s[:, :, 3:6], s[:, :, 6:7])
    loss = ref.flow_arbitrary.compute_l2_error(pred, b4["space_samples_tgt"])
    loss.backward()
    gold["train_arb_loss"] = np.array(loss.item(), np.float64)
    gold["train_arb_pred"] = pred.detach().numpy()
    gold["train_arb_gradnorms"] = np.array([float(p.grad.norm()) if p.grad is not None else -1.0
                                            for _, p in m.named_parameters()], np.float64)
    sd = m.state_dict()
    gold["train_arb_buf::model_canonicalize.encoder.transformer_begin.bn.running_mean"] = \
        sd["model_canonicalize.encoder.transformer_begin.bn.running_mean"].numpy().copy()
    gold["train_arb_buf::model_canonicalize.encoder.transformer_begin.bn.num_batches_tracked"] = \
        sd["model_canonicalize.encoder.transformer_begin.bn.num_batches_tracked"].numpy().copy()

    np.savez_compressed(os.path.join(OUT, "tdnet_reference.npz"), **gold)
    print({k: v.shape for k, v in gold.items()})


if __name__ == "__main__":
    main()
