"""Interpolation decoder (reference: model/decoder/interpolation_decoder.py:8-88; the reference's ablation decoder,
`decoder: interp`): Gaussian kernel regression of the anchor features at every query, then a ResNet-FC stack.

Same constructor arguments, sub-module names (state_dict keys) and call signature as the reference. The kernel
regression is written so that nothing of size [B, Q, N, 3] is materialised (the reference expands the anchors per query,
interpolation_decoder.py:57): squared distances come from one batched GEMM-shaped expansion, the normalised weights
[B, Q, N] multiply the anchor features in one bmm.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from nsdp_b200.model.decoder.blocks import ResnetBlockFC


class PointInterpDecoder(nn.Module):
    def __init__(self, dim_inp, dim, out_dim=3, hidden_dim=50, n_blocks=5):
        super().__init__()
        self.n_blocks = n_blocks
        self.fc0 = nn.Linear(dim_inp, dim)
        self.fc1 = nn.Linear(dim, hidden_dim)
        self.blocks = nn.ModuleList([ResnetBlockFC(hidden_dim) for _ in range(n_blocks)])
        self.fc_c = nn.ModuleList([nn.Linear(dim, hidden_dim) for _ in range(n_blocks)])
        self.fc_out = nn.Linear(hidden_dim, out_dim)
        self.actvn = F.relu
        self.var = 0.2 ** 2

    def sample_point_feature(self, q, p, fea):
        """q (B,M,3) queries, p (B,N,3) anchors, fea (B,N,c) -> (B,M,c): weights exp(-(|p - q| + 1e-5)^2 / var), normalised
        over the anchors (interpolation_decoder.py:50-66)."""
        d = torch.cdist(q, p, p=2.0, compute_mode="donot_use_mm_for_euclid_dist")   # exact differences, as the reference
        weight = (-((d + 10e-6) ** 2) / self.var).exp()
        weight = weight / weight.sum(dim=2, keepdim=True)
        return weight @ fea

    def forward(self, xyz_q, encoding):
        xyz, feats = encoding["anchors"], encoding["anchor_feats"]
        lat_rep = self.fc0(self.sample_point_feature(xyz_q, xyz, feats))
        net = self.fc1(F.relu(lat_rep))
        for i in range(self.n_blocks):
            net = net + self.fc_c[i](lat_rep)
            net = self.blocks[i](net)
        return self.fc_out(self.actvn(net))
