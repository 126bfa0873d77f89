"""Arbitrary-pose composition (reference: model/flow_arbitrary.py:7-85): source -> canonical with the
backward TDNet (for the space samples AND the surface samples), then canonical -> target with the forward
TDNet whose encoder input is cat[surface_src2cano, surface_tgt, mask].

Inference (test_on_batch_with_arbitrary) encodes each distinct encoder input once: 2 encoder passes where the reference
runs 6 (flow_arbitrary.py:65-85).

The reference runs the canonicalise ENCODER twice on identical input (flow_arbitrary.py:19-20). Here it is
encoded once and decoded for both query sets; to stay bit-compatible with the reference's BatchNorm
bookkeeping in train() mode (every BN updates its running stats twice per step) the single
pass runs with the equivalent momentum 2m - m^2 and bumps num_batches_tracked twice.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from nsdp_b200 import dist as nsdp_dist
from nsdp_b200.graph import graphed_train_step
from nsdp_b200.model.utils import compute_l2_error


def _bn_layers(module):
    return [m for m in module.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm)]


class FlowArbitrary(nn.Module):
    def __init__(self, cfg, model_canonicalize, model_deform):
        super().__init__()
        self.model_canonicalize = model_canonicalize
        self.model_deform = model_deform

    def _canonicalize_twice(self, space_samples_src, surface_samples_src):
        cano = self.model_canonicalize
        bns = _bn_layers(cano.encoder) if self.training else []
        # Two identical passes update every running stat twice: r2 = r0 + (2m - m^2)(batch - r0). One pass with
        # the momentum temporarily set to 2m - m^2 gives the same buffers (and num_batches_tracked += 2) without
        # touching tensors autograd has saved.
        saved = [bn.momentum for bn in bns]
        for bn in bns:
            if bn.momentum is not None:
                bn.momentum = 2.0 * bn.momentum - bn.momentum * bn.momentum
        try:
            encoding = cano.encode(surface_samples_src)
        finally:
            for bn, m in zip(bns, saved):
                bn.momentum = m
        with torch.no_grad():
            for bn in bns:
                if bn.num_batches_tracked is not None:
                    bn.num_batches_tracked += 1
        space = cano.decode(space_samples_src, encoding)
        surface = cano.decode(surface_samples_src, encoding)
        return space, surface

    def forward(self, space_samples_src, surface_samples_src, surface_samples_tgt, cano_handle_sample_mask):
        space_src2cano, surface_src2cano = self._canonicalize_twice(space_samples_src, surface_samples_src)
        deform_inputs = torch.cat([surface_src2cano, surface_samples_tgt, cano_handle_sample_mask], dim=-1).contiguous()
        return self.model_deform(space_src2cano, deform_inputs)


def _split_inputs(data_dict):
    s = data_dict["surface_samples_inputs"]
    return s[:, :, 0:3], s[:, :, 3:6], s[:, :, 6:7]


def _train_step_with_arbitrary(model, optimizer, data_dict):
    nsdp_dist.zero_grad(model, optimizer)
    src, tgt, mask = _split_inputs(data_dict)
    pred = model(data_dict["space_samples_src"], src, tgt, mask)
    loss = compute_l2_error(pred, data_dict["space_samples_tgt"])
    loss.backward()
    return loss


def _finish_step(model, optimizer):
    nsdp_dist.allreduce_gradients(model)
    optimizer.step()


def train_on_batch_with_arbitrary(model, optimizer, data_dict, config):
    """flow_arbitrary.py:30-48; captured into a CUDA graph after a few calls (nsdp_b200/graph.py)."""
    return graphed_train_step(model, optimizer, data_dict, _train_step_with_arbitrary, _finish_step,
                              keys=("surface_samples_inputs", "space_samples_src", "space_samples_tgt"))


@torch.no_grad()
def validate_on_batch_with_arbitrary(model, data_dict, config):
    src, tgt, mask = _split_inputs(data_dict)
    pred = model(data_dict["space_samples_src"], src, tgt, mask)
    return compute_l2_error(pred, data_dict["space_samples_tgt"]).item()


@torch.no_grad()
def test_on_batch_with_arbitrary(model, data_dict, config, compute_loss=False):
    src, tgt, mask = _split_inputs(data_dict)
    if model.training:
        # train-mode BatchNorm: every encoder pass updates running stats, keep the reference's passes
        data_dict["surface_samples_tgt_pred"] = model(src, src, tgt, mask)
        verts_pred = model(data_dict["verts_src"], src, tgt, mask)
    else:
        # The reference's two model(...) calls (flow_arbitrary.py:71-79) encode the same source surface with the
        # canonicaliser 2 x 2 times and the same cat[surface_src2cano, tgt, mask] with the deform net twice: in eval mode
        # all of that is identical work. Encode each once (2 encoder passes instead of 6), then decode per query set; the
        # surface samples' canonical positions double as the first query set's.
        cano, deform = model.model_canonicalize, model.model_deform
        cano_enc = cano.encode(src)
        surface_src2cano = cano.decode(src.contiguous(), cano_enc)
        deform_enc = deform.encode(torch.cat([surface_src2cano, tgt, mask], dim=-1).contiguous())
        data_dict["surface_samples_tgt_pred"] = deform.decode(surface_src2cano, deform_enc)
        verts_pred = deform.decode(cano.decode(data_dict["verts_src"], cano_enc), deform_enc)
    data_dict["verts_tgt_pred"] = verts_pred
    if compute_loss:
        loss = compute_l2_error(verts_pred, data_dict["verts_tgt"])
    else:
        loss = torch.zeros((1), dtype=torch.float32)
    return loss.item(), data_dict
