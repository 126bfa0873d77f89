"""TDNet glue (reference: model/deformation_networks.py:12-109): encoder(surface) -> decoder(queries).

`Deformation_Networks` and the three *_on_batch_with_cano functions keep the reference names, arguments,
data_dict keys and return values, so train.py / test.py / run.py call them unchanged. Additions that do not
change single-GPU semantics: the data-parallel gradient all-reduce between backward() and step() when a
process group is initialised (nsdp_b200.dist), and encode-once caching in test_on_batch (the reference runs
the identical encoder twice, deformation_networks.py:96-101).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from nsdp_b200 import dist as nsdp_dist
from nsdp_b200.graph import graphed_forward, graphed_train_step
from nsdp_b200.model.decoder import decoder_dict
from nsdp_b200.model.encoder import encoder_dict
from nsdp_b200.model.utils import compute_l2_error


class Deformation_Networks(nn.Module):
    def __init__(self, cfg, no_input_corr=False):
        super().__init__()
        self.no_input_corr = no_input_corr
        use_normals = cfg["model"]["use_normals"]
        # inp_feat_dim rules of the reference (deformation_networks.py:16-30)
        if no_input_corr:
            has_features, inp_feat_dim = (True, 3) if use_normals else (False, 0)
        else:
            has_features, inp_feat_dim = (True, 7) if use_normals else (True, 4)
        mcfg = cfg["model"]
        self.encoder = encoder_dict[mcfg["encoder"]](has_features=has_features, inp_feat_dim=inp_feat_dim,
                                                      **mcfg["encoder_kwargs"])
        self.decoder = decoder_dict[mcfg["decoder"]](**mcfg["decoder_kwargs"])

    def encode(self, surface_samples_inputs):
        if self.no_input_corr:
            return self.encoder(surface_samples_inputs[:, :, 0:3].contiguous())
        return self.encoder(surface_samples_inputs)

    def decode(self, points, encoding):
        return self.decoder(points, encoding)

    def _forward(self, points, surface_samples_inputs):
        return self.decode(points, self.encode(surface_samples_inputs))

    def forward(self, points, surface_samples_inputs):
        # eval mode + no_grad + CUDA inputs of a shape seen before: replayed as one CUDA graph (nsdp_b200/graph.py)
        return graphed_forward(self, (points, surface_samples_inputs), self._forward)


def _train_step_with_cano(model, optimizer, data_dict):
    nsdp_dist.zero_grad(model, optimizer)
    pred = model(data_dict["space_samples_src"], data_dict["surface_samples_inputs"])
    loss = compute_l2_error(pred, data_dict["space_samples_tgt"])
    loss.backward()
    return loss


def _finish_step(model, optimizer):
    nsdp_dist.allreduce_gradients(model)
    optimizer.step()


def train_on_batch_with_cano(model, optimizer, data_dict, config):
    """deformation_networks.py:63-77. The step itself is `_train_step_with_cano`; on a GPU it is captured into a CUDA graph
    after a few calls and replayed (nsdp_b200/graph.py) — same arithmetic, one launch per step."""
    return graphed_train_step(model, optimizer, data_dict, _train_step_with_cano, _finish_step,
                              keys=("surface_samples_inputs", "space_samples_src", "space_samples_tgt"))


@torch.no_grad()
def validate_on_batch_with_cano(model, data_dict, config):
    pred = model(data_dict["space_samples_src"], data_dict["surface_samples_inputs"])
    return compute_l2_error(pred, data_dict["space_samples_tgt"]).item()


@torch.no_grad()
def test_on_batch_with_cano(model, data_dict, config, compute_loss=False):
    surface = data_dict["surface_samples_inputs"]
    if model.training:
        # train-mode BatchNorm: every encoder pass updates running stats, keep the reference's two passes
        data_dict["surface_samples_tgt_pred"] = model(data_dict["surface_samples_src"], surface)
        verts_pred = model(data_dict["verts_src"], surface)
    else:
        encoding = model.encode(surface)  # identical for both query sets in eval mode
        data_dict["surface_samples_tgt_pred"] = model.decode(data_dict["surface_samples_src"], encoding)
        verts_pred = model.decode(data_dict["verts_src"], encoding)
    data_dict["verts_tgt_pred"] = verts_pred
    if compute_loss:
        loss = compute_l2_error(verts_pred, data_dict["verts_tgt"])
    else:
        loss = torch.zeros((1), dtype=torch.float32)
    return loss.item(), data_dict
