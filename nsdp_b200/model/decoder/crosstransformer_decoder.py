"""Neural-field decoder (reference: model/decoder/crosstransformer_decoder.py:6-70) on two fused kernels:
ops.vector_attention (anchor cross-attention, with the global token) -> ops.resnet_tail (init_enc, 5 x (fc_c +
ResnetBlockFC), fc_out). Activations between the 17 linear layers never touch HBM."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from nsdp_b200 import ops
from nsdp_b200.model.decoder.blocks import CrossTransformerBlock, ResnetBlockFC


class CrossTransformerDecoder(nn.Module):
    def __init__(self, dim_inp, dim, nneigh=7, hidden_dim=64, n_blocks=5, out_dim=1):
        super().__init__()
        self.dim = dim
        self.n_blocks = n_blocks
        self.ct1 = CrossTransformerBlock(dim_inp, dim, nneigh=nneigh)
        self.init_enc = nn.Linear(dim, hidden_dim)
        self.blocks = nn.ModuleList([ResnetBlockFC(hidden_dim) for _ in range(n_blocks)])
        self.fc_c = nn.ModuleList([nn.Linear(dim, hidden_dim) for _ in range(n_blocks)])
        self.fc_out = nn.Linear(hidden_dim, out_dim)
        self.actvn = F.relu

    def packed_tail_weights(self):
        """K-major, concatenated weights in the layout nsdp_tail_args documents."""
        wc = torch.cat([self.init_enc.weight] + [l.weight for l in self.fc_c], dim=0)       # ((1+n)H, C)
        bc = torch.cat([self.init_enc.bias] + [l.bias for l in self.fc_c], dim=0)
        w0 = torch.stack([b.fc_0.weight.t() for b in self.blocks]).contiguous()              # (n, H, H)
        b0 = torch.stack([b.fc_0.bias for b in self.blocks]).contiguous()
        w1 = torch.stack([b.fc_1.weight.t() for b in self.blocks]).contiguous()
        b1 = torch.stack([b.fc_1.bias for b in self.blocks]).contiguous()
        return (wc.t().contiguous(), bc.contiguous(), w0, b0, w1, b1, self.fc_out.weight.t().contiguous(),
                self.fc_out.bias.contiguous())

    def forward(self, xyz_q, encoding):
        lat = self.ct1(xyz_q, encoding["z"], encoding["anchors"], encoding["anchor_feats"])   # (B, Q, dim)
        B, Q, C = lat.shape
        out = ops.resnet_tail(lat.reshape(B * Q, C), *self.packed_tail_weights())
        return out.reshape(B, Q, -1)
