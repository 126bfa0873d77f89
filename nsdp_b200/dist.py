"""Data parallelism over shapes: one process per GPU, weights replicated, every rank takes a
contiguous slice of the batch, and ONE flat all-reduce of the gradients per step.

The reference has no distributed code at all (device hard-wired to cuda:0, train.py:74-77). This module is
what sits behind the unchanged train_on_batch_* functions: they call allreduce_gradients(model) between
backward() and optimizer.step(); it is a no-op unless a process group exists.

BatchNorm semantics: by default local per-rank batch statistics (what DistributedDataParallel would give the
reference). The all-reduce averages gradients so that the mean-loss semantics match a single-process batch.
Optional `syncbn` mode (NSDP_B200_SYNCBN=1 or convert_sync_batchnorm(model)): every BatchNorm1d reduces
(sum x, sum x^2, count) over all ranks, which reproduces the single-process batch-32 numbers of SURVEY.md §8e
(outputs, gradients and running statistics) at the cost of two tiny collectives per layer and step.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as td

_FLAT = {}  # id(model) -> (flat buffer, [params])


def is_active() -> bool:
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def init_process_group(backend: Optional[str] = None) -> None:
    """Create the default process group from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun)."""
    if td.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    td.init_process_group(backend=backend)


def maybe_init_from_env(model, device) -> None:
    """Called by build_model: joins the job when launched under torchrun / nsdp_b200.launch, and makes the
    replicas start from identical weights (rank 0 broadcasts)."""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1 or os.environ.get("NSDP_B200_DP", "1") == "0":
        return
    init_process_group()
    broadcast_parameters(model)
    if os.environ.get("NSDP_B200_SYNCBN", "0") == "1":
        convert_sync_batchnorm(model)


@torch.no_grad()
def broadcast_parameters(model, src: int = 0) -> None:
    if not is_active():
        return
    for t in list(model.parameters()) + list(model.buffers()):
        td.broadcast(t.data, src=src)


def _flat_for(model) -> tuple:
    key = id(model)
    params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
    entry = _FLAT.get(key)
    n = sum(p.numel() for p in params)
    if entry is None or entry[0].numel() != n or entry[0].device != params[0].device:
        flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
        _FLAT[key] = entry = (flat, params)
    return entry


@torch.no_grad()
def allreduce_gradients(model) -> None:
    """Average gradients across ranks with a single collective over one flat fp32 buffer (17.97 MB for a
    TDNet, 35.94 MB for FlowArbitrary). Parameters that received no gradient (the unused q/k/v weights of the
    pos_only block) contribute zeros, so every rank reduces the same layout."""
    if not is_active():
        return
    flat, params = _flat_for(model)
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            flat[off:off + n].zero_()
        else:
            flat[off:off + n].copy_(p.grad.reshape(-1))
        off += n
    td.all_reduce(flat, op=td.ReduceOp.SUM)
    flat.div_(td.get_world_size())
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].reshape(p.shape).clone()
        else:
            p.grad.copy_(flat[off:off + n].reshape(p.shape))
        off += n


def shard_batch(data_dict: dict, rank: Optional[int] = None, world: Optional[int] = None) -> dict:
    """Contiguous batch slice of every tensor for this rank: shapes [r*B/W, (r+1)*B/W)."""
    if rank is None:
        rank = td.get_rank() if is_active() else 0
    if world is None:
        world = td.get_world_size() if is_active() else 1
    out = {}
    for k, v in data_dict.items():
        if torch.is_tensor(v) and v.dim() > 0:
            B = v.shape[0]
            if B % world != 0:
                raise ValueError(f"batch {B} of '{k}' is not divisible by world size {world}")
            per = B // world
            out[k] = v[rank * per:(rank + 1) * per]
        else:
            out[k] = v
    return out


@torch.no_grad()
def sharded_decode(decode_fn, points: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """Inference-time query sharding: the decoder is independent per query point, so with the
    (cheap, replicated) encoding computed on every rank, rank r decodes the contiguous query slice
    [r*Q/W, (r+1)*Q/W) of `points` (B, Q, 3) and the slices are all-gathered back into (B, Q, C) on every rank — meshes
    with millions of vertices (test.py:127-151, run.py:121-139) decode W times faster. `decode_fn(points_slice)` is
    e.g. `lambda p: model.decode(p, encoding)`. Without a process group this is just `decode_fn(points)`."""
    if rank is None:
        rank = td.get_rank() if is_active() else 0
    if world is None:
        world = td.get_world_size() if is_active() else 1
    if world == 1:
        return decode_fn(points)
    Q = points.shape[1]
    per = (Q + world - 1) // world                        # equal-sized slices: the last one is padded with its first query
    lo, hi = min(rank * per, Q), min((rank + 1) * per, Q)
    mine = points[:, lo:hi]
    if mine.shape[1] < per:
        pad = (mine[:, :1] if mine.shape[1] else points[:, :1]).expand(-1, per - mine.shape[1], -1)
        mine = torch.cat([mine, pad], dim=1)
    out = decode_fn(mine.contiguous()).contiguous()
    parts = [torch.empty_like(out) for _ in range(world)]
    td.all_gather(parts, out)
    return torch.cat(parts, dim=1)[:, :Q]


# ---------------------------------------------------------------------------------------------------
# optional: BatchNorm statistics over the GLOBAL batch (SURVEY.md §8e "syncbn")
# ---------------------------------------------------------------------------------------------------
class _SyncBatchNormFn(torch.autograd.Function):
    """y = (x - mean) * invstd * weight + bias with mean / var over the rows of ALL ranks. x is (R, C)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps):
        C = x.shape[1]
        packed = torch.empty(2 * C + 1, dtype=torch.float32, device=x.device)
        packed[:C] = x.sum(0)
        packed[C:2 * C] = (x * x).sum(0)
        packed[2 * C] = x.shape[0]
        td.all_reduce(packed, op=td.ReduceOp.SUM)
        n = packed[2 * C]
        mean = packed[:C] / n
        var = (packed[C:2 * C] / n - mean * mean).clamp_min_(0.0)
        invstd = torch.rsqrt(var + eps)
        if running_mean is not None:
            with torch.no_grad():
                running_mean.mul_(1 - momentum).add_(momentum * mean)
                running_var.mul_(1 - momentum).add_(momentum * var * (n / (n - 1)))   # unbiased, like nn.BatchNorm1d
        xhat = (x - mean) * invstd
        ctx.save_for_backward(xhat, weight, invstd, n)
        return xhat * weight + bias

    @staticmethod
    def backward(ctx, dy):
        xhat, weight, invstd, n = ctx.saved_tensors
        C = xhat.shape[1]
        sums = torch.cat([dy.sum(0), (dy * xhat).sum(0)])
        d_bias, d_weight = sums[:C].clone(), sums[C:].clone()          # parameter gradients stay LOCAL sums; the flat gradient
        td.all_reduce(sums, op=td.ReduceOp.SUM)                        # all-reduce averages them like every other parameter
        dx = (weight * invstd) * (dy - sums[:C] / n - xhat * (sums[C:] / n))
        return dx, d_weight, d_bias, None, None, None, None


class SyncBatchNorm1d(torch.nn.BatchNorm1d):
    """nn.BatchNorm1d (same parameters, buffers and state_dict keys) whose training-mode statistics span all ranks.
    Falls back to the parent's behaviour in eval mode and when no process group is active."""

    def forward(self, x):
        if not (self.training and is_active()) or x.dim() != 2:
            return super().forward(x)
        if self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
        return _SyncBatchNormFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var,
                                      self.momentum, self.eps)


def convert_sync_batchnorm(model):
    """In-place: every nn.BatchNorm1d of `model` becomes a SyncBatchNorm1d (the mirror applies BN to (B*n, C) rows, see
    model/encoder/blocks.py `_bn_rows`). state_dict keys and values are untouched."""
    for m in model.modules():
        if type(m) is torch.nn.BatchNorm1d:
            m.__class__ = SyncBatchNorm1d
    return model
