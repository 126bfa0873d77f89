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
import weakref
from typing import Iterable, List, Optional

import torch
import torch.distributed as td



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


class _GradBuckets:
    """All gradients of a model as views of ONE flat fp32 buffer, laid out in REVERSE parameter order (roughly the order
    in which backward() finalises them: decoder first, early encoder layers last) and cut into contiguous buckets of
    >= `bucket_bytes`. A post-accumulate hook per parameter counts a bucket down; as soon as a bucket (and every bucket
    before it: all ranks must issue collectives in the same order) is complete its slice is all-reduced asynchronously
    — NCCL runs it on its own stream, overlapped with the rest of the backward — and `finish()` only waits.

    No per-parameter copies into or out of the flat buffer, no `div_` (ReduceOp.AVG on NCCL; one scale per bucket on
    gloo). Parameters that never receive a gradient (the unused q/k/v weights of the pos_only block) keep zeros in their
    views; which ones do is learnt on the first step, during which every bucket is reduced at the end."""

    def __init__(self, model, bucket_bytes: int = 4 << 20):
        self.params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad][::-1]
        p0 = self.params[0]
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=p0.dtype, device=p0.device)
        self.views, self.bucket_of, self.bounds = [], [], [0]
        off, cur = 0, 0
        for p in self.params:
            n = p.numel()
            self.views.append(self.flat[off:off + n].view_as(p))
            self.bucket_of.append(len(self.bounds) - 1)
            off += n
            cur += n * p.element_size()
            if cur >= bucket_bytes:
                self.bounds.append(off)
                cur = 0
        if self.bounds[-1] != off:
            self.bounds.append(off)
        self.nb = len(self.bounds) - 1
        self.index = {id(p): i for i, p in enumerate(self.params)}
        self.expected = None                       # per bucket: the parameters whose hooks fire (learnt on step 1)
        self.remaining = None
        self.fired_ids = set()
        self.launched = 0                          # buckets [0, launched) are in flight / done this step
        self.works, self.late = [], []
        self.synced_state = False
        self.avg = td.get_backend() == "nccl"
        self.overlap = os.environ.get("NSDP_B200_OVERLAP", "1") != "0"
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    def attach(self) -> None:
        for p, v in zip(self.params, self.views):
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def zero(self) -> None:
        self.attach()
        self.flat.zero_()
        self.fired_ids = set()
        self.remaining = [set(s) for s in self.expected] if (self.expected is not None and self.overlap) else None
        self.launched = 0
        self.works, self.late = [], []

    def _launch(self, b: int) -> None:
        t = self.flat[self.bounds[b]:self.bounds[b + 1]]
        if self.avg:
            self.works.append((td.all_reduce(t, op=td.ReduceOp.AVG, async_op=True), None))
        else:
            self.works.append((td.all_reduce(t, op=td.ReduceOp.SUM, async_op=True), t))

    def _hook(self, p) -> None:
        i = self.index.get(id(p))
        if i is None or not is_active():
            return
        if p.grad.data_ptr() != self.views[i].data_ptr():
            self.views[i].copy_(p.grad)            # somebody replaced .grad (zero_grad(set_to_none=True)): fold it back
            p.grad = self.views[i]
        b = self.bucket_of[i]
        self.fired_ids.add(i)
        if b < self.launched:                      # arrived while / after its bucket was being reduced
            self.late.append(i)
            return
        if self.remaining is None:
            return
        self.remaining[b].discard(i)
        while self.launched < self.nb and not self.remaining[self.launched]:
            self._launch(self.launched)
            self.launched += 1

    def finish(self) -> None:
        if self.launched == 0 and not self.overlap:
            # nothing left during backward (CUDA-graph replay, or overlap switched off): ONE collective over the whole buffer
            if self.avg:
                self.works.append((td.all_reduce(self.flat, op=td.ReduceOp.AVG, async_op=True), None))
            else:
                self.works.append((td.all_reduce(self.flat, op=td.ReduceOp.SUM, async_op=True), self.flat))
            self.launched = self.nb
        while self.launched < self.nb:
            self._launch(self.launched)
            self.launched += 1
        world = td.get_world_size()
        for w, t in self.works:
            w.wait()
            if t is not None:
                t.mul_(1.0 / world)
        late = self.late
        if self.fired_ids and (self.expected is None or late):
            self.expected = [set(i for i in self.fired_ids if self.bucket_of[i] == b) for b in range(self.nb)]
        # ready for the next step even if nobody calls zero() in Python (CUDA-graph replays run no Python in between)
        self.works, self.late, self.launched, self.fired_ids = [], [], 0, set()
        if late:
            # its in-place accumulation raced with the collective in flight: this step's gradient of that bucket is not
            # trustworthy. Only possible when a parameter that had no gradient on the first step gets one later.
            raise RuntimeError("gradients of parameters %s arrived after their bucket had been all-reduced (the set of "
                               "parameters with gradients changed between steps); set NSDP_B200_OVERLAP=0" % sorted(set(late)))


_BUCKETS = weakref.WeakKeyDictionary()  # model -> _GradBuckets (dies with the model: an id() can be reused)


def _buckets_for(model) -> _GradBuckets:
    bk = _BUCKETS.get(model)
    if bk is None:
        bk = _BUCKETS[model] = _GradBuckets(model, int(os.environ.get("NSDP_B200_BUCKET_BYTES", 4 << 20)))
    return bk


@torch.no_grad()
def sync_training_state(model, optimizer=None) -> None:
    """Rank 0's parameters, buffers and optimizer state become everybody's. build_model() broadcasts the initial weights,
    but an unchanged train.py loads checkpoints AFTER build_model (train.py:152-156): called on the first train step so
    that replicas that resumed from different files (or did not resume) cannot drift apart silently."""
    if not is_active():
        return
    broadcast_parameters(model)
    if optimizer is None:
        return
    dev = next(model.parameters()).device
    # the replicas must agree on the STRUCTURE of the optimizer state before tensors can be broadcast
    n_state = torch.tensor([sum(len(s) for s in optimizer.state.values())], dtype=torch.int64, device=dev)
    lo, hi = n_state.clone(), n_state.clone()
    td.all_reduce(lo, op=td.ReduceOp.MIN)
    td.all_reduce(hi, op=td.ReduceOp.MAX)
    if int(lo) != int(hi):
        raise RuntimeError("data-parallel replicas resumed from different checkpoints (optimizer state present on some "
                           "ranks only): every rank must read rank 0's model_*/opt_* files (nsdp_b200.launch does this)")
    for group in optimizer.param_groups:
        for p in group["params"]:
            for k, v in sorted(optimizer.state.get(p, {}).items()):
                if torch.is_tensor(v):
                    t = v.to(dev) if v.device != dev else v
                    td.broadcast(t, src=0)
                    if t is not v:
                        v.copy_(t)


def set_overlap(model, on: bool) -> None:
    """Turn the hook-driven early bucket launches off (CUDA-graph capture: a hook would fire only while capturing)."""
    _buckets_for(model).overlap = bool(on)


def zero_grad(model, optimizer) -> None:
    """optimizer.zero_grad() of the reference's train_on_batch_* (deformation_networks.py:65). Under data parallelism the
    gradients live in one flat buffer (one memset) and keep their storage from step to step."""
    if not is_active():
        optimizer.zero_grad()
        return
    bk = _buckets_for(model)
    if not bk.synced_state:
        sync_training_state(model, optimizer)
        bk.synced_state = True
    bk.zero()


@torch.no_grad()
def allreduce_gradients(model) -> None:
    """Average the gradients across ranks (17.97 MB for a TDNet, 35.94 MB for FlowArbitrary). When the step went through
    zero_grad() above, most of the traffic already left during backward(); this waits for it and reduces what is left.
    Parameters that received no gradient contribute zeros, so every rank reduces the same layout. A model whose
    gradients were produced without zero_grad() (plain `.backward()` on fresh `.grad`s) is folded into the flat buffer
    first."""
    if not is_active():
        return
    bk = _buckets_for(model)
    for p, v in zip(bk.params, bk.views):
        if p.grad is None:
            p.grad = v
        elif p.grad.data_ptr() != v.data_ptr():
            v.copy_(p.grad)
            p.grad = v
    bk.finish()


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
        packed = torch.empty(2 * C + 1, dtype=x.dtype if x.dtype == torch.float64 else torch.float32, device=x.device)
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
