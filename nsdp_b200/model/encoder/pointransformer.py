"""Point-transformer encoder (reference: model/encoder/pointransformer.py:6-140) on the nsdp_b200 blocks.

Constructor arguments, sub-module names and the returned dict ({'z', 'anchors', 'anchor_feats'}) are the
reference's; see blocks.py for what runs underneath.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from nsdp_b200 import ops
from nsdp_b200.model.encoder.blocks import ElementwiseMLP, TransformerBlock, TransitionDown, fold_sites


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

    def _fold_all(self):
        """Kernel-ready weights of every attention site of the encoder in one batched pass (blocks.fold_sites)."""
        keys, sites = ["begin"], [self.transformer_begin.fold_site()]
        for level, down in enumerate(self.transition_downs):
            if hasattr(down.sa, "fold_site"):                       # attentive set abstraction: two sites
                keys += [("sa", level, 0), ("sa", level, 1)]
                sites += down.sa.fold_site()
            keys.append(("td", level))
            sites.append(self.transformer_downs[level].fold_site())
        for i, blk in enumerate(self.final_transformers):
            keys.append(("final", i))
            sites.append(blk.fold_site())
        folded = dict(zip(keys, fold_sites(sites)))
        for level in range(len(self.transition_downs)):
            if ("sa", level, 0) in folded:
                folded[("sa", level)] = (folded.pop(("sa", level, 0)), folded.pop(("sa", level, 1)))
        return folded

    def forward(self, xyz, intermediate_out_path=None):
        if intermediate_out_path is not None:
            raise NotImplementedError("intermediate point-cloud dumps are a debugging aid of the reference "
                                      "(pointransformer.py:94-136) and are not part of the hot path")
        folds = self._fold_all()
        if self.has_features:
            raw = xyz[:, :, 3:]
            feats = ops.linear(raw, self.enc_sdf.weight, self.enc_sdf.bias)     # nn.Linear(4 -> 120) on B * N rows
            xyz = xyz[:, :, :3].contiguous()
            feats = self.transformer_begin(xyz, feats, feats_from=(raw, self.enc_sdf), folded=folds["begin"])
        else:
            feats = self.transformer_begin(xyz, folded=folds["begin"])

        for level, down in enumerate(self.transition_downs):
            xyz, feats = down(xyz, feats, folded=folds.get(("sa", level)))
            feats = self.elementwise_extras[level](feats)
            feats = self.transformer_downs[level](xyz, feats, folded=folds[("td", level)])
            if level == 0 and self.d_reduced != self.d_transformer:
                feats = self.fc1(feats)
            feats = self.elementwise[level](feats)

        for i, (block, mlp) in enumerate(zip(self.final_transformers, self.final_elementwise)):
            feats = mlp(block(xyz, feats, folded=folds[("final", i)]))

        z = self.fc_middle(feats.max(dim=1)[0])
        return {"z": z, "anchors": xyz, "anchor_feats": feats}
