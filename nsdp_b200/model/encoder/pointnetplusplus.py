"""PointNet++-style encoder (reference: model/encoder/pointnetplusplus.py:5-96; the reference's ablation encoder,
`encoder: pointnet++`) on the nsdp_b200 blocks.

Same constructor arguments, sub-module names (hence state_dict keys) and returned dict as the reference. What runs
underneath: FPS and k-NN of every max-pool set abstraction are the sm_100a kernels (blocks.PointNetSetAbstraction),
the final full self-attention blocks are the fused vector-attention kernel (blocks.TransformerBlock, group_all).
"""
from __future__ import annotations

import torch.nn as nn

from nsdp_b200.model.encoder.blocks import ElementwiseMLP, TransformerBlock, TransitionDown


class PointNetPlusPlusEncoder(nn.Module):
    def __init__(self, npoints_per_layer, nneighbor, d_transformer, nfinal_transformers, has_features=False,
                 inp_feat_dim=1):
        super().__init__()
        self.d_transformer = d_transformer
        self.has_features = has_features
        self.inp_feat_dim = inp_feat_dim
        self.fc_middle = nn.Sequential(nn.Linear(d_transformer, d_transformer), nn.ReLU(),
                                       nn.Linear(d_transformer, d_transformer))
        # per-point lifting of the features (or of the coordinates when there are none), pointnetplusplus.py:35-46
        self.fc_begin = nn.Sequential(nn.Linear(inp_feat_dim if has_features else 3, d_transformer), nn.ReLU(),
                                      nn.Linear(d_transformer, d_transformer))
        self.transition_downs = nn.ModuleList()
        self.elementwise = nn.ModuleList()
        for level in range(len(npoints_per_layer) - 1):
            n_in, n_out = npoints_per_layer[level], npoints_per_layer[level + 1]
            self.transition_downs.append(TransitionDown(n_out, min(nneighbor, n_in), d_transformer, type="maxpool"))
            self.elementwise.append(ElementwiseMLP(d_transformer))
        self.final_transformers = nn.ModuleList(
            [TransformerBlock(d_transformer, -1, group_all=True) for _ in range(nfinal_transformers)])
        self.final_elementwise = nn.ModuleList([ElementwiseMLP(dim=d_transformer) for _ in range(nfinal_transformers)])

    def forward(self, xyz):
        if self.has_features:
            feats = self.fc_begin(xyz[:, :, 3:].contiguous())
            xyz = xyz[:, :, 0:3].contiguous()
        else:
            feats = self.fc_begin(xyz)
        for down, mlp in zip(self.transition_downs, self.elementwise):
            xyz, feats = down(xyz, feats)
            feats = mlp(feats)
        for block, mlp in zip(self.final_transformers, self.final_elementwise):
            feats = mlp(block(xyz, feats))
        return {"z": self.fc_middle(feats.max(dim=1)[0]), "anchors": xyz, "anchor_feats": feats}
