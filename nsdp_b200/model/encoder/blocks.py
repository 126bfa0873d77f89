"""Point-transformer encoder blocks on the nsdp_b200 kernels.

Same class names, constructor arguments, attribute names (hence state_dict keys) and forward signatures as
the reference's model/encoder/blocks.py, so reference checkpoints load unchanged. What differs is HOW the
forward runs: the pair-level chain (gather -> fc_delta -> fc_gamma -> softmax -> weighted sum,
blocks.py:104-126 and :290-308) is ONE fused CUDA kernel (ops.vector_attention); k-NN is the tiled top-k
kernel instead of a full distance matrix + argsort (blocks.py:101-102, 287-288); FPS is the
cluster-resident kernel (blocks.py:283). Only per-point linear layers and BatchNorm stay torch ops
(plain cuBLAS GEMMs / cuDNN-free batch_norm on (B*n, C) views).

Algebra used to shrink the pair-level work (exact in real arithmetic, fp32 rounding differs ~1e-7):
    fc_gamma[0](q_i - k_j + delta_ij) = (Wg0 Wq) x_i - (Wg0 Wk) x_j + (Wg0 Wd2) h_ij + (Wg0 bd2 + bg0)
so the kernel receives per-POINT tables Q' and K' and one folded matrix W' = Wg0 Wd2; fc_gamma[2].bias
is constant over the softmax axis and drops out. All folds are differentiable torch ops, so autograd
delivers the gradients of the original parameters.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from nsdp_b200 import ops
from nsdp_b200.model.utils import index_points, knn_indices


def _mlp2(d_in: int, d: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(d_in, d), nn.ReLU(), nn.Linear(d, d))


def _bn_rows(bn: nn.BatchNorm1d, x: torch.Tensor) -> torch.Tensor:
    """BatchNorm1d over the channel dim of a (B, n, C) tensor == the reference's permute -> bn -> permute."""
    B, n, C = x.shape
    return bn(x.reshape(B * n, C)).reshape(B, n, C)


def _pointwise(conv: nn.Conv1d, x: torch.Tensor) -> torch.Tensor:
    """A kernel-size-1 Conv1d applied to (B, n, C) rows."""
    return F.linear(x, conv.weight.squeeze(-1), conv.bias)


def _fused_linear(x: torch.Tensor, weights):
    """[F.linear(x, w) for w in weights] as ONE GEMM against the row-concatenated weights (the per-point projections of a
    block share their input; three or four [B*n, d] x [d, d] launches become one, forward and backward), returned as
    contiguous tensors (the kernels take dense (B, n, d) tables)."""
    outs = F.linear(x, torch.cat(list(weights), dim=0)).split([w.shape[0] for w in weights], dim=-1)
    return [o.contiguous() for o in outs]


def _fused_linear_through(x_in: torch.Tensor, lin: nn.Linear, weights):
    """[F.linear(lin(x_in), w) for w in weights] without the wide intermediate GEMM: the projections of features that are
    themselves a Linear of a NARROW input (the encoder's first block: feats = enc_sdf(4 input channels), 32 768 rows) fold
    into one [rows, c_in] x [c_in, sum d_out] product — (W lin.weight) x + W lin.bias — 30x fewer FLOPs than projecting the
    120-wide features (exact in real arithmetic; autograd carries the gradients back through the fold)."""
    wcat = torch.cat(list(weights), dim=0)
    outs = ops.linear(x_in, wcat @ lin.weight, torch.mv(wcat, lin.bias)).split([w.shape[0] for w in weights], dim=-1)
    return [o.contiguous() for o in outs]


def fold_pair_mlps(fc_delta: nn.Sequential, fc_gamma: nn.Sequential):
    """Kernel-side weights for one (fc_delta, fc_gamma) pair: see nsdp_vattn_args."""
    wd0, bd0 = fc_delta[0].weight, fc_delta[0].bias
    wd2, bd2 = fc_delta[2].weight, fc_delta[2].bias
    wg0, bg0 = fc_gamma[0].weight, fc_gamma[0].bias
    wg2 = fc_gamma[2].weight
    wp = wg0 @ wd2
    return dict(
        wd0=wd0.contiguous(), bd0=bd0.contiguous(),
        wd2t=wd2.t().contiguous(),
        wpt=wp.t().contiguous(),
        wg2t=wg2.t().contiguous(),
        pc=torch.mv(wg0, bd2) + bg0,
        vc=bd2.contiguous(),
        # the same matrices un-transposed, for the backward's data-gradient products (no gradient flows through these)
        wd2n=wd2.detach().contiguous(), wpn=wp.detach(), wg2n=wg2.detach().contiguous(),
    )


def fold_sites(sites):
    """fold_pair_mlps + the `Wg0 @ W` projection folds for MANY attention sites at once. `sites` is a list of
    (fc_delta, fc_gamma, [Linear, ...]): the Linears are the per-point projections that feed fc_gamma[0] (w_qs, w_ks, ...).
    Sites of equal width are stacked and folded with batched products: the encoder has 10 sites of 2 widths, and folding
    them one by one costs ~13 tiny launches per site forward and twice that backward (weights change every step, so the folds
    are part of every step). Returns one (operands dict, [folded projection weights]) pair per site, equal to what
    fold_pair_mlps / `wg0 @ w` give (same arithmetic, batched)."""
    out = [None] * len(sites)
    groups = {}
    for i, (fd, fg, projs) in enumerate(sites):
        groups.setdefault(fd[2].weight.shape[0], []).append(i)
    for d, members in groups.items():
        if len(members) == 1:
            i = members[0]
            fd, fg, projs = sites[i]
            out[i] = (fold_pair_mlps(fd, fg), [fg[0].weight @ p.weight for p in projs])
            continue
        wd2 = torch.stack([sites[i][0][2].weight for i in members])            # (S, d, d)
        bd2 = torch.stack([sites[i][0][2].bias for i in members])              # (S, d)
        wg0 = torch.stack([sites[i][1][0].weight for i in members])
        bg0 = torch.stack([sites[i][1][0].bias for i in members])
        wg2 = torch.stack([sites[i][1][2].weight for i in members])
        wd2t = wd2.transpose(1, 2).contiguous()
        wp = torch.bmm(wg0, wd2)
        wpt = wp.transpose(1, 2).contiguous()
        wg2t = wg2.transpose(1, 2).contiguous()
        wd2n, wpn, wg2n = wd2.detach(), wp.detach(), wg2.detach()
        pc = torch.bmm(wg0, bd2.unsqueeze(-1)).squeeze(-1) + bg0
        # projection folds: every (site, projection) pair is one batch entry
        owner = [(s, p) for s, i in enumerate(members) for p in sites[i][2]]
        # (no index tensors: a Python-list index would be a host -> device copy, which a CUDA-graph capture refuses)
        folded = torch.bmm(torch.stack([wg0[s] for s, _ in owner]), torch.stack([p.weight for _, p in owner])) if owner else None
        per_site = {s: [] for s in range(len(members))}
        for n, (s, _) in enumerate(owner):
            per_site[s].append(folded[n])
        for s, i in enumerate(members):
            fd = sites[i][0]
            out[i] = (dict(wd0=fd[0].weight.contiguous(), bd0=fd[0].bias.contiguous(), wd2t=wd2t[s], wpt=wpt[s], wg2t=wg2t[s],
                           pc=pc[s], vc=bd2[s], wd2n=wd2n[s], wpn=wpn[s], wg2n=wg2n[s]), per_site[s])
    return out


class TransformerBlock(nn.Module):
    """Local / global vector self-attention (reference: model/encoder/blocks.py:52-134)."""

    def __init__(self, d_model, k, pos_only=False, group_all=False) -> None:
        super().__init__()
        self.pos_only = pos_only
        self.bn = nn.BatchNorm1d(d_model)
        self.fc_delta = _mlp2(3, d_model)
        self.fc_gamma = _mlp2(d_model, d_model)
        self.w_qs = nn.Linear(d_model, d_model, bias=False)
        self.w_ks = nn.Linear(d_model, d_model, bias=False)
        self.w_vs = nn.Linear(d_model, d_model, bias=False)
        self.k = k
        self.group_all = group_all

    def fold_site(self):
        """This block's entry for fold_sites()."""
        return (self.fc_delta, self.fc_gamma, [] if self.pos_only else [self.w_qs, self.w_ks])

    def forward(self, xyz, feats=None, feats_from=None, folded=None):
        """Optimisation hints of the mirror (the reference signature is (xyz, feats)): `feats_from = (x_in, lin)`, the caller's
        statement that feats == lin(x_in) with a narrow x_in (see _fused_linear_through); `folded`, this block's entry of
        fold_sites() computed ahead of time together with the other blocks'."""
        B, n, _ = xyz.shape
        xyz = xyz.contiguous()
        idx = None if self.group_all else knn_indices(xyz, xyz, min(self.k, n))
        w, fp = folded if folded is not None else fold_sites([self.fold_site()])[0]
        if self.pos_only:
            res = ops.vector_attention(xyz, xyz, idx, None, None, None, sign=1.0, **w)
        else:
            proj = (fp[0], fp[1], self.w_vs.weight)
            if feats_from is not None and feats_from[0].shape[-1] * 4 <= feats.shape[-1]:
                qp, kp, vp = _fused_linear_through(feats_from[0], feats_from[1], proj)
            else:
                qp, kp, vp = _fused_linear(feats, proj)
            res = ops.vector_attention(xyz, xyz, idx, qp, kp, vp, sign=1.0, **w) + feats
        return _bn_rows(self.bn, res)


class ElementwiseMLP(nn.Module):
    """bn3(x + relu(bn2(conv2(relu(bn1(conv1 x)))))) (reference: model/encoder/blocks.py:137-159)."""

    def __init__(self, dim):
        super().__init__()
        self.conv1 = nn.Conv1d(dim, dim, 1)
        self.bn1 = nn.BatchNorm1d(dim)
        self.conv2 = nn.Conv1d(dim, dim, 1)
        self.bn2 = nn.BatchNorm1d(dim)
        self.bn3 = nn.BatchNorm1d(dim)

    def forward(self, x):
        if self.training and type(self.bn1) is not nn.BatchNorm1d:
            # optional syncbn mode (nsdp_b200.dist.convert_sync_batchnorm): the statistics span all ranks, one collective per
            # BatchNorm, so the three layers run one by one
            h = F.relu(_bn_rows(self.bn1, _pointwise(self.conv1, x)))
            h = F.relu(_bn_rows(self.bn2, _pointwise(self.conv2, h)))
            return _bn_rows(self.bn3, x + h)
        # one fused op: 4 kernels forward / 7 backward (csrc/emlp.cu) instead of the ~12 / ~30 cuDNN + elementwise launches
        return ops.elementwise_mlp(x, self.conv1, self.bn1, self.conv2, self.bn2, self.bn3)


class TransformerSetAbstraction(nn.Module):
    """FPS down-sampling + two stacked cross vector-attentions (reference: model/encoder/blocks.py:221-314)."""

    def __init__(self, npoint, nneigh, dim):
        super().__init__()
        self.npoint = npoint
        self.nneigh = nneigh
        self.bnorm0 = nn.BatchNorm1d(dim)
        self.bnorm1 = nn.BatchNorm1d(dim)
        self.bnorm2 = nn.BatchNorm1d(dim)
        self.bn1 = nn.BatchNorm1d(dim)
        self.conv1 = nn.Conv1d(dim, dim, 1)
        self.conv2 = nn.Conv1d(dim, dim, 1)
        self.fc_delta1 = _mlp2(3, dim)
        self.fc_gamma1 = _mlp2(dim, dim)
        self.fc_gamma2 = _mlp2(dim, dim)
        self.w_qs = nn.Linear(dim, dim, bias=False)
        self.w_ks = nn.Linear(dim, dim, bias=False)
        self.w_vs = nn.Linear(dim, dim, bias=False)
        self.w_qs2 = nn.Linear(dim, dim, bias=False)
        self.w_ks2 = nn.Linear(dim, dim, bias=False)
        self.w_vs2 = nn.Linear(dim, dim, bias=False)

    def fold_site(self):
        """Two entries for fold_sites(): (fc_delta1, fc_gamma1) and (fc_delta1, fc_gamma2)."""
        return [(self.fc_delta1, self.fc_gamma1, [self.w_qs, self.w_ks]), (self.fc_delta1, self.fc_gamma2, [self.w_qs2, self.w_ks2])]

    def forward(self, xyz, points, folded=None):
        xyz = xyz.contiguous()
        N = xyz.shape[1]
        with torch.no_grad():
            fps_idx = ops.furthest_point_sampling(xyz.detach(), self.npoint)
            new_xyz = index_points(xyz.detach(), fps_idx).contiguous()  # detached, as in blocks.py:282-285
            idx = knn_indices(new_xyz, xyz, min(self.nneigh, N))
        centre = index_points(points, fps_idx)

        (w1, (fq1, fk1)), (w2, (fq2, fk2)) = folded if folded is not None else fold_sites(self.fold_site())
        qp = F.linear(centre, fq1)
        # the four projections of the full cloud (both attention stages) share their input: one GEMM
        kp, vp, kp2, vp2 = _fused_linear(points, (fk1, self.w_vs.weight, fk2, self.w_vs2.weight))
        # rel = neighbour - centre (blocks.py:295) -> sign = -1
        res1 = ops.vector_attention(new_xyz, xyz, idx, qp, kp, vp, sign=-1.0, **w1)
        res1 = res1 + _pointwise(self.conv2, F.relu(_bn_rows(self.bn1, _pointwise(self.conv1, res1))))
        res1 = _bn_rows(self.bnorm0, res1)

        qp2 = F.linear(res1, fq2)          # second stage: same delta MLP, second gamma MLP
        res2 = ops.vector_attention(new_xyz, xyz, idx, qp2, kp2, vp2, sign=-1.0, **w2)

        out = _bn_rows(self.bnorm1, res1 + res2) + centre
        return new_xyz, _bn_rows(self.bnorm2, out)


class PointNetSetAbstraction(nn.Module):
    """PointNet++-style max-pool set abstraction (reference: model/encoder/blocks.py:162-217; ablation only —
    no shipped config selects it). FPS and k-NN run on the nsdp_b200 kernels; the rest is per-point torch ops."""

    def __init__(self, npoint, nneigh, in_channel, dim):
        super().__init__()
        self.npoint = npoint
        self.nneigh = nneigh
        self.fc1 = nn.Linear(in_channel, dim)
        self.conv1 = nn.Conv1d(dim, dim, 1)
        self.conv2 = nn.Conv1d(dim, dim, 1)
        self.bn1 = nn.BatchNorm1d(dim)
        self.bn2 = nn.BatchNorm1d(dim)
        self.bn = nn.BatchNorm1d(dim)

    def forward(self, xyz, points):
        xyz = xyz.contiguous()
        with torch.no_grad():
            fps_idx = ops.furthest_point_sampling(xyz.detach(), self.npoint)
        new_xyz = index_points(xyz, fps_idx)
        points = self.fc1(points)
        centre = index_points(points, fps_idx)
        h = F.relu(_bn_rows(self.bn1, _pointwise(self.conv1, points)))
        points = points + F.relu(_bn_rows(self.bn2, _pointwise(self.conv2, h)))
        idx = knn_indices(new_xyz, xyz, min(self.nneigh, xyz.shape[1]))
        pooled = index_points(points, idx).max(dim=2)[0]
        return new_xyz, _bn_rows(self.bn, centre + pooled)


class TransitionDown(nn.Module):
    """Wrapper choosing the set-abstraction flavour (reference: model/encoder/blocks.py:18-49)."""

    def __init__(self, npoint, nneighbor, dim, type="attentive") -> None:
        super().__init__()
        if type == "attentive":
            self.sa = TransformerSetAbstraction(npoint, nneighbor, dim)
        elif type == "maxpool":
            self.sa = PointNetSetAbstraction(npoint, nneighbor, dim, dim)
        else:
            raise ValueError("Set Abstraction type " + type + " unknown!")

    def forward(self, xyz, feats, folded=None):
        return self.sa(xyz, feats) if folded is None else self.sa(xyz, feats, folded=folded)
