"""Data parallelism over shapes: one process per GPU, weights replicated, every rank takes a
contiguous slice of the batch, and ONE flat all-reduce of the gradients per step.

The reference has no distributed code at all (device hard-wired to cuda:0, train.py:74-77). This module is
what sits behind the unchanged train_on_batch_* functions: they call allreduce_gradients(model) between
backward() and optimizer.step(); it is a no-op unless a process group exists.

BatchNorm semantics: local per-rank batch statistics (what DistributedDataParallel would give the
reference). The all-reduce averages gradients so that the mean-loss semantics match a single-process batch.
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
