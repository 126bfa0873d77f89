"""Point-transformer encoder (reference: model/encoder/pointransformer.py:6-140) on the nsdp_b200 blocks.

Constructor arguments, sub-module names and the returned dict ({'z', 'anchors', 'anchor_feats'}) are the
reference's; see blocks.py for what runs underneath.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from nsdp_b200.model.encoder.blocks import ElementwiseMLP, TransformerBlock, TransitionDown


class PointTransformerEncoder(nn.Module):
    def __init__(self, npoints_per_layer, nneighbor, nneighbor_reduced, nfinal_transformers, d_transformer, d_reduced,
                 full_SA=False, has_features=False, inp_feat_dim=1):
        super().__init__()
        self.d_reduced = d_reduced
        self.d_transformer = d_transformer
        self.has_features = has_features

        self.fc_middle = nn.Sequential(nn.Linear(d_transformer, d_transformer), nn.ReLU(),
                                       nn.Linear(d_transformer, d_transformer))
        if has_features:
            self.enc_sdf = nn.Linear(inp_feat_dim, d_reduced)
        self.transformer_begin = TransformerBlock(d_reduced, nneighbor_reduced, pos_only=not has_features)
        self.transition_downs = nn.ModuleList()
        self.transformer_downs = nn.ModuleList()
        self.elementwise = nn.ModuleList()
        self.elementwise_extras = nn.ModuleList()
        if d_reduced != d_transformer:
            self.fc1 = nn.Linear(d_reduced, d_transformer)

        for level in range(len(npoints_per_layer) - 1):
            n_in, n_out = npoints_per_layer[level], npoints_per_layer[level + 1]
            dim = d_reduced if level == 0 else d_transformer
            # k is clamped with the CONFIGURED cardinalities (pointransformer.py:63-67), not the actual N
            self.transition_downs.append(TransitionDown(n_out, min(nneighbor, n_in), dim))
            self.elementwise_extras.append(ElementwiseMLP(dim))
            self.transformer_downs.append(TransformerBlock(dim, min(nneighbor, n_out)))
            self.elementwise.append(ElementwiseMLP(d_transformer))

        self.final_transformers = nn.ModuleList(
            [TransformerBlock(d_transformer, 2 * nneighbor, group_all=full_SA) for _ in range(nfinal_transformers)])
        self.final_elementwise = nn.ModuleList([ElementwiseMLP(dim=d_transformer) for _ in range(nfinal_transformers)])

    def forward(self, xyz, intermediate_out_path=None):
        if intermediate_out_path is not None:
            raise NotImplementedError("intermediate point-cloud dumps are a debugging aid of the reference "
                                      "(pointransformer.py:94-136) and are not part of the hot path")
        if self.has_features:
            raw = xyz[:, :, 3:]
            feats = self.enc_sdf(raw)
            xyz = xyz[:, :, :3].contiguous()
            feats = self.transformer_begin(xyz, feats, feats_from=(raw, self.enc_sdf))
        else:
            feats = self.transformer_begin(xyz)

        for level, down in enumerate(self.transition_downs):
            xyz, feats = down(xyz, feats)
            feats = self.elementwise_extras[level](feats)
            feats = self.transformer_downs[level](xyz, feats)
            if level == 0 and self.d_reduced != self.d_transformer:
                feats = self.fc1(feats)
            feats = self.elementwise[level](feats)

        for block, mlp in zip(self.final_transformers, self.final_elementwise):
            feats = mlp(block(xyz, feats))

        z = self.fc_middle(feats.max(dim=1)[0])
        return {"z": z, "anchors": xyz, "anchor_feats": feats}
