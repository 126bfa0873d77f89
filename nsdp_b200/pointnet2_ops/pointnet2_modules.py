"""PointNet++ set-abstraction / feature-propagation modules with the reference's public names, constructor arguments,
tensor layouts and state_dict keys (pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py:9-209), on the nsdp_b200
kernels: FPS + gather for the centroids, ball query + grouping (or group-all) per scale, 3-NN + inverse-distance
interpolation for propagation. `pointnet2_ops/__init__.py:1` imports this module, so the drop-in package needs it even
though TDNet itself only calls furthest_point_sample.

The per-scale "shared MLP" stays an nn.Sequential of 1x1 Conv2d / BatchNorm2d / ReLU exactly as the reference builds it
(keys `mlps.<scale>.<layer>.weight`, ...), so reference checkpoints of these modules load unchanged.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from nsdp_b200.pointnet2_ops import pointnet2_utils as pu


def build_shared_mlp(mlp_spec: Sequence[int], bn: bool = True) -> nn.Sequential:
    """[c0, c1, ..., cn] -> n x (Conv2d 1x1 (bias only without BN), [BatchNorm2d], ReLU); pointnet2_modules.py:9-19."""
    stack: List[nn.Module] = []
    for c_in, c_out in zip(mlp_spec[:-1], mlp_spec[1:]):
        stack.append(nn.Conv2d(c_in, c_out, kernel_size=1, bias=not bn))
        if bn:
            stack.append(nn.BatchNorm2d(c_out))
        stack.append(nn.ReLU(True))
    return nn.Sequential(*stack)


class _PointnetSAModuleBase(nn.Module):
    """xyz (B,N,3), features (B,C,N) or None -> new_xyz (B,npoint,3) (None when grouping all), (B, sum_k mlp_k[-1], npoint)."""

    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def _centroids(self, xyz: torch.Tensor) -> Optional[torch.Tensor]:
        if self.npoint is None:
            return None
        picks = pu.furthest_point_sample(xyz, self.npoint)                          # (B, npoint) int32, bit-exact FPS
        return pu.gather_operation(xyz.transpose(1, 2).contiguous(), picks).transpose(1, 2).contiguous()

    def forward(self, xyz: torch.Tensor, features: Optional[torch.Tensor]) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        new_xyz = self._centroids(xyz)
        pooled = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            grouped = mlp(grouper(xyz, new_xyz, features))                          # (B, mlp[-1], npoint, nsample)
            # max over the neighbourhood; max_pool2d (not amax) so that ties route their gradient exactly as in the reference
            pooled.append(F.max_pool2d(grouped, kernel_size=[1, grouped.size(3)]).squeeze(-1))
        return new_xyz, torch.cat(pooled, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Multi-scale grouping: one (radius, nsample, mlp) triple per scale; npoint=None groups the whole cloud.
    Like the reference, `use_xyz` bumps mlp[0] by 3 IN PLACE in the caller's list (pointnet2_modules.py:114-116)."""

    def __init__(self, npoint, radii, nsamples, mlps, bn=True, use_xyz=True):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            self.groupers.append(pu.QueryAndGroup(radius, nsample, use_xyz=use_xyz) if npoint is not None
                                 else pu.GroupAll(use_xyz))
            if use_xyz:
                spec[0] += 3
            self.mlps.append(build_shared_mlp(spec, bn))


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction (pointnet2_modules.py:121-150)."""

    def __init__(self, mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz)


class PointnetFPModule(nn.Module):
    """Feature propagation: unknown (B,n,3), known (B,m,3) or None, unknow_feats (B,C1,n) or None, known_feats (B,C2,m)
    -> (B, mlp[-1], n); pointnet2_modules.py:153-209."""

    def __init__(self, mlp, bn=True):
        super().__init__()
        self.mlp = build_shared_mlp(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        if known is None:
            carried = known_feats.expand(*known_feats.shape[:2], unknown.size(1))
        else:
            dist, idx = pu.three_nn(unknown, known)
            inv = 1.0 / (dist + 1e-8)
            carried = pu.three_interpolate(known_feats, idx, inv / inv.sum(dim=2, keepdim=True))
        stacked = carried if unknow_feats is None else torch.cat([carried, unknow_feats], dim=1)
        return self.mlp(stacked.unsqueeze(-1)).squeeze(-1)
