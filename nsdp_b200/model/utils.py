"""Mirror of the reference's model/utils.py helpers that sit on the hot path.

`index_points` / `compute_l2_error` keep the reference signatures (model/utils.py:8-11, 58-70);
`knn_indices` replaces the `square_distance(...).argsort()[:, :, :k]` idiom (model/utils.py:39-55) with the
tiled top-k kernel, so the [B, M, N] matrix is never built.
"""
from __future__ import annotations

import torch

from nsdp_b200 import ops


def compute_l2_error(points_pred: torch.Tensor, points_gt: torch.Tensor) -> torch.Tensor:
    """mean over (b, q) of 0.5 * ||pred - gt||^2  (model/utils.py:8-11)."""
    return ((points_pred - points_gt).pow(2).sum(dim=2) * 0.5).mean()


def index_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """points (B,N,C), idx (B,S[,K]) int32/int64 -> (B,S[,K],C)  (model/utils.py:58-70)."""
    shape = idx.shape
    flat = idx.reshape(shape[0], -1).long()
    out = torch.gather(points, 1, flat.unsqueeze(-1).expand(-1, -1, points.shape[-1]))
    return out.reshape(*shape, points.shape[-1])


@torch.no_grad()
def knn_indices(query: torch.Tensor, ref: torch.Tensor, k: int) -> torch.Tensor:
    """(B,M,3), (B,N,3) -> (B,M,k) int32, ascending (distance, index)."""
    return ops.knn(query.detach().contiguous(), ref.detach().contiguous(), k)


def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """Kept for API compatibility (model/utils.py:39-55); the hot path uses knn_indices instead."""
    return torch.sum((src[:, :, None] - dst[:, None]) ** 2, dim=-1)


def farthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """The reference's torch-level FPS (model/utils.py:73-93), which its encoder keeps as a commented-out alternative to the
    CUDA operator (model/encoder/blocks.py:283-284): RANDOM start index per shape, int64 indices (B, npoint), no skip rule.
    Kept for API compatibility and off the hot path — the model calls ops.furthest_point_sampling (start 0, the
    `pointnet2_ops` semantics). Runs on whatever device `xyz` lives on."""
    B, N, _ = xyz.shape
    picks = torch.zeros(B, npoint, dtype=torch.long, device=xyz.device)
    nearest = torch.full((B, N), 1e10, dtype=xyz.dtype, device=xyz.device)     # squared distance to the closest pick so far
    current = torch.randint(0, N, (B,), dtype=torch.long).to(xyz.device)
    rows = torch.arange(B, dtype=torch.long, device=xyz.device)
    for i in range(npoint):
        picks[:, i] = current
        d = ((xyz - xyz[rows, current].view(B, 1, 3)) ** 2).sum(-1)
        nearest = torch.where(d < nearest, d, nearest)
        current = nearest.max(dim=-1)[1]
    return picks


def fibonacci_sphere(samples: int = 1):
    """`samples` points spread evenly over the unit sphere by the golden-angle spiral, as a (samples, 3) float64 numpy array
    (model/utils.py:13-36)."""
    import math

    import numpy as np
    i = np.arange(samples, dtype=np.float64)
    y = 1.0 - (i / float(samples - 1)) * 2.0 if samples > 1 else np.ones(1)
    radius = np.sqrt(np.maximum(1.0 - y * y, 0.0))
    theta = math.pi * (3.0 - math.sqrt(5.0)) * i
    return np.stack([np.cos(theta) * radius, y, np.sin(theta) * radius], axis=1)
