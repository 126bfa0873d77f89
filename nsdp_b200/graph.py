"""Whole-step CUDA graphs behind the unchanged train_on_batch_* API.

A training step of this model is ~1100 kernel launches, most of them microsecond-sized encoder glue; issued one by one
from Python + autograd they leave the GPU idle for milliseconds per step. All shapes are static from step to step
(fixed batch, fixed point counts), there is no host synchronisation between `zero_grad` and `loss.item()`, and every
nsdp_b200 kernel is launched on the current stream through the C ABI — so the WHOLE step (forward, loss, backward, fused
Adam) is captured once into a CUDA graph and replayed: one launch per step.

    loss = graphed_train_step(model, optimizer, data_dict, forward_backward, finish)     # -> float

`forward_backward(model, optimizer, data_dict) -> loss tensor` is zero_grad + forward + loss + backward;
`finish(model, optimizer)` is the gradient all-reduce (a no-op on one GPU) + optimizer.step(). On one GPU both are captured
into ONE graph. Under data parallelism the graph holds forward_backward only — the gradients land in the flat bucket
buffer of nsdp_b200.dist — and `finish` runs eagerly after the replay (one NCCL all-reduce + the fused Adam kernels).

Protocol per (model, optimizer, input shapes): the first WARMUP calls run eagerly on a side stream (lazy initialisation:
cuBLAS handles, kernel attributes, gradient buffers, Adam state), the next call captures, later calls copy the batch into
the graph's static input buffers and replay. Anything that prevents capture (non-capturable optimizer, a failed capture,
...) falls back to the eager path, loudly, once.
The learning rate is read from the param groups on every call; a change (model/learningrate.py adjust_learning_rate,
train.py:188) re-captures, because a Python float lr is baked into the captured launch.
NSDP_B200_GRAPH=0 disables the whole mechanism.
"""
from __future__ import annotations

import os
import sys
import weakref

import torch

from nsdp_b200 import ops

ENABLED = os.environ.get("NSDP_B200_GRAPH", "1") != "0"
WARMUP = 3
_STATE = weakref.WeakKeyDictionary()      # model -> {signature: _Entry}


class _Entry:
    __slots__ = ("calls", "graph", "static", "loss", "lrs", "launches", "failed", "opt_graph", "opt_ref")

    def __init__(self, optimizer=None):
        self.calls, self.graph, self.static, self.loss = 0, None, None, None
        self.lrs, self.launches, self.failed, self.opt_graph = None, 0, False, None
        # the signature holds id(optimizer), and an id can be reused by a NEW optimizer once the old one is collected: the
        # captured launches would then keep updating the dead optimizer's state tensors
        self.opt_ref = weakref.ref(optimizer) if optimizer is not None else None

    def stale(self, optimizer) -> bool:
        return self.opt_ref is not None and self.opt_ref() is not optimizer


def _signature(model, optimizer, data_dict, keys):
    # what the captured launches depend on besides the values in the static buffers: input shapes, train / eval mode, which
    # parameters are trainable (freezing a sub-network changes the autograd graph)
    sig = [id(optimizer), model.training, sum(1 for p in model.parameters() if p.requires_grad)]
    for k in keys:
        v = data_dict[k]
        sig.append((k, tuple(v.shape), v.dtype, str(v.device)))
    return tuple(sig)


def _lrs(optimizer):
    out = []
    for g in optimizer.param_groups:
        lr = g["lr"]
        out.append(float(lr) if not torch.is_tensor(lr) else None)      # a tensor lr lives on the device: no re-capture
        out.append(float(g.get("weight_decay", 0.0)))
    return tuple(out)


def _capturable(optimizer) -> bool:
    return isinstance(optimizer, torch.optim.Adam) and all(g.get("capturable", False) for g in optimizer.param_groups)


def _usable(model, optimizer, data_dict, keys) -> bool:
    if not ENABLED or ops.TIMING:
        return False
    from nsdp_b200 import dist
    if not dist.is_active() and not _capturable(optimizer):
        return False
    if dist.is_active() and any(isinstance(m, dist.SyncBatchNorm1d) for m in model.modules()):
        return False          # syncbn mode issues collectives inside forward / backward: stays eager
    tensors = [data_dict.get(k) for k in keys]
    return bool(tensors) and all(torch.is_tensor(v) and v.is_cuda for v in tensors) and not torch.cuda.is_current_stream_capturing()


def graphed_train_step(model, optimizer, data_dict, forward_backward, finish, keys):
    """`keys`: the entries of data_dict the step reads (the datasets put ~20 tensors into a sample, the step uses 3): only these
    get static buffers inside the graph; nothing else of data_dict reaches the captured code."""
    from nsdp_b200 import dist
    split = dist.is_active()          # data parallel: the collective and the optimizer stay outside the graph

    def eager_fn(model, optimizer, data_dict):
        loss = forward_backward(model, optimizer, data_dict)
        finish(model, optimizer)
        return loss

    if not _usable(model, optimizer, data_dict, keys):
        return eager_fn(model, optimizer, data_dict).item()
    per_model = _STATE.setdefault(model, {})
    sig = _signature(model, optimizer, data_dict, keys)
    e = per_model.get(sig)
    if e is None or e.stale(optimizer):
        e = per_model[sig] = _Entry(optimizer)
    if e.failed:
        return eager_fn(model, optimizer, data_dict).item()
    lrs = _lrs(optimizer)
    if e.graph is not None and e.lrs != lrs:
        e.graph, e.static, e.loss, e.opt_graph = None, None, None, None           # learning rate changed: capture again
    if e.graph is None:
        if e.calls < WARMUP:
            e.calls += 1
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                loss = eager_fn(model, optimizer, data_dict)
            torch.cuda.current_stream().wait_stream(side)
            return loss.item()
        try:
            if split:
                dist.set_overlap(model, False)       # no collectives from autograd hooks while capturing / replaying
                _capture(e, model, optimizer, data_dict, forward_backward, keys)
                if _capturable(optimizer):           # the fused Adam launches as a second, tiny graph after the collective
                    og = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(og):
                        optimizer.step()
                    e.opt_graph = og
            else:
                _capture(e, model, optimizer, data_dict, eager_fn, keys)
            e.lrs = lrs
        except Exception as exc:   # noqa: BLE001 — never lose a training step over an optimisation
            e.failed = True
            e.graph = None
            torch.cuda.synchronize()
            print(f"[nsdp_b200.graph] CUDA-graph capture of the training step failed ({type(exc).__name__}: {exc}); "
                  "continuing on the eager path", file=sys.stderr)
            return eager_fn(model, optimizer, data_dict).item()
    for k, buf in e.static.items():
        src = data_dict[k]
        if src.data_ptr() != buf.data_ptr():
            buf.copy_(src, non_blocking=True)
    e.graph.replay()
    ops._count(e.launches)
    if split:
        if e.opt_graph is not None:
            dist.allreduce_gradients(model)
            e.opt_graph.replay()
        else:
            finish(model, optimizer)
    return e.loss.item()


def _capture(e: _Entry, model, optimizer, data_dict, eager_fn, keys) -> None:
    static = {k: data_dict[k].clone() for k in keys}
    passthrough = {}
    # gradients must be (re)created INSIDE the capture so that they live in the graph's memory pool (single GPU); under
    # data parallelism they are views of the flat bucket buffer, which exists since the first warm-up step
    from nsdp_b200 import dist
    if not dist.is_active():
        optimizer.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    before = ops.LAUNCHES
    # the parameters' AccumulateGrad nodes were created on the warm-up stream; inside the capture everything runs on the
    # capture stream, which is what is wanted here
    warn = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
    if warn is not None:
        warn(False)
    with torch.cuda.graph(g):
        loss = eager_fn(model, optimizer, {**static, **passthrough})
    e.launches = ops.LAUNCHES - before
    ops.LAUNCHES = before            # the capture pass executed nothing; replays account for their kernel calls
    e.graph, e.static, e.loss = g, static, loss
    # the capture itself executed nothing: the caller replays right away with the current batch


# ---------------------------------------------------------------------------------------------------------------------
# inference: eval-mode forward under no_grad (validate_on_batch, test.py / run.py on fixed-size inputs, bench --forward-only)
# ---------------------------------------------------------------------------------------------------------------------
_FWD_STATE = weakref.WeakKeyDictionary()      # model -> {signature: _Entry}
MAX_FORWARD_GRAPHS = 4                        # per model: meshes of many different sizes (test.py) simply stay eager


def graphed_forward(model, inputs, eager_fn):
    """eager_fn(*inputs) -> tensor, for a model in eval mode with gradients disabled: no side effects, so after WARMUP eager
    calls per input-shape signature the forward is captured and replayed (the result is copied out of the graph's static
    output buffer). Everything else runs eager_fn directly."""
    if (not ENABLED or ops.TIMING or model.training or torch.is_grad_enabled()
            or not all(torch.is_tensor(t) and t.is_cuda for t in inputs) or torch.cuda.is_current_stream_capturing()):
        return eager_fn(*inputs)
    per_model = _FWD_STATE.setdefault(model, {})
    sig = tuple((tuple(t.shape), t.dtype, str(t.device)) for t in inputs)
    e = per_model.get(sig)
    if e is None:
        if len(per_model) >= MAX_FORWARD_GRAPHS:
            return eager_fn(*inputs)
        e = per_model[sig] = _Entry()
    if e.failed:
        return eager_fn(*inputs)
    if e.graph is None:
        if e.calls < WARMUP:
            e.calls += 1
            return eager_fn(*inputs)
        try:
            static = [t.clone() for t in inputs]
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            before = ops.LAUNCHES
            with torch.cuda.graph(graph):
                out = eager_fn(*static)
            e.launches = ops.LAUNCHES - before
            ops.LAUNCHES = before
            e.graph, e.static, e.loss = graph, static, out
        except Exception as exc:   # noqa: BLE001
            e.failed, e.graph = True, None
            torch.cuda.synchronize()
            print(f"[nsdp_b200.graph] CUDA-graph capture of the eval forward failed ({type(exc).__name__}: {exc}); "
                  "continuing on the eager path", file=sys.stderr)
            return eager_fn(*inputs)
    for buf, src in zip(e.static, inputs):
        if src.data_ptr() != buf.data_ptr():
            buf.copy_(src, non_blocking=True)
    e.graph.replay()
    ops._count(e.launches)
    return e.loss.clone()
