"""CUDA-graph execution behind the unchanged API (nsdp_b200/graph.py): a graph-replayed training run must reproduce the
eager run step for step — with a different batch every step (static input buffers are refreshed), across a learning-rate
change (re-capture) — and the graph-replayed eval forward must equal the eager one."""
import copy

import numpy as np
import pytest
import torch

from nsdp_b200 import graph, ops, synth
from nsdp_b200.model import build_model, optimizer_factory

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(schemas, mtype="forward"):
    cfg = synth.make_config(mtype)
    model, train_on_batch, _, _ = build_model(cfg, device=DEV)
    model.load_state_dict(synth.named_state_dict([(k, s) for k, s in schemas[mtype]], seed=0))
    model.train()
    _, opt = optimizer_factory(cfg["training"], model.parameters())
    return cfg, model, train_on_batch, opt


def _batches(n, B=2, N=600, Q=500):
    return [{k: v.to(DEV) for k, v in synth.forward_batch(B, N, Q, seed=100 + i, fp16_grid=False).items()} for i in range(n)]


@torch.no_grad()
def _copy_state(dst_model, dst_opt, src_model, src_opt):
    """dst <- src, IN PLACE (a captured graph keeps pointing at the tensors it was captured with)."""
    for d, s in zip(list(dst_model.parameters()) + list(dst_model.buffers()), list(src_model.parameters()) + list(src_model.buffers())):
        d.copy_(s)
    for dp, sp in zip(dst_opt.param_groups[0]["params"], src_opt.param_groups[0]["params"]):
        ds, ss = dst_opt.state.get(dp, {}), src_opt.state.get(sp, {})
        assert set(ds) == set(ss)
        for k in ss:
            if torch.is_tensor(ss[k]):
                ds[k].copy_(ss[k])


def test_graph_replayed_training_matches_eager(schemas, monkeypatch):
    """Model G trains through the graph path, model E eagerly; before every step E is reset to G's state, so each step is
    compared from IDENTICAL weights / Adam state (two free-running fp32 trajectories drift apart chaotically — atomics'
    summation order, amplified by Adam's sign-like early updates — whether or not a graph is involved)."""
    batches = _batches(8)
    cfg, model_g, train_g, opt_g = _setup(schemas)
    _, model_e, train_e, opt_e = _setup(schemas)
    before = ops.LAUNCHES
    per_step = None
    for i, b in enumerate(batches):
        if i == 6:                                   # train.py:188 adjust_learning_rate: a Python float in the param group
            for g in opt_g.param_groups + opt_e.param_groups:
                g["lr"] = 1e-4
        if i > 0:
            _copy_state(model_e, opt_e, model_g, opt_g)
        monkeypatch.setattr(graph, "ENABLED", True)
        n0 = ops.LAUNCHES
        loss_g = train_g(model_g, opt_g, dict(b), cfg)
        n_g = ops.LAUNCHES - n0
        monkeypatch.setattr(graph, "ENABLED", False)
        n0 = ops.LAUNCHES
        loss_e = train_e(model_e, opt_e, dict(b), cfg)
        n_e = ops.LAUNCHES - n0
        assert n_g == n_e > 0                         # a replay accounts for the kernel calls it contains
        assert abs(loss_g - loss_e) <= 1e-5 * abs(loss_e) + 1e-9, (i, loss_g, loss_e)
        for (k, pg), pe in zip(model_g.state_dict().items(), model_e.state_dict().values()):
            if pg.is_floating_point():
                err = float((pg - pe).norm() / pe.norm().clamp_min(1e-12))
                # one Adam step from identical state. Parameters whose true gradient is 0 (a bias feeding a batch-statistics
                # BatchNorm) see pure rounding noise, and Adam turns its random SIGN into a full +-lr step: 2 lr / |p| ~ 1e-2
                assert err < 2e-2, (i, k, err)
            else:
                assert torch.equal(pg, pe), (i, k)
    (e,) = graph._STATE[model_g].values()
    assert e.graph is not None and not e.failed and e.lrs[0] == 1e-4      # captured at step 4, re-captured after the lr change
    assert model_e not in graph._STATE or all(x.graph is None for x in graph._STATE[model_e].values())


def test_graph_replayed_eval_forward_matches_eager(schemas, monkeypatch):
    cfg, model, _, _ = _setup(schemas)
    model.eval()
    batches = _batches(6)
    monkeypatch.setattr(graph, "ENABLED", False)
    with torch.no_grad():
        want = [model(b["space_samples_src"], b["surface_samples_inputs"]).cpu() for b in batches]
    monkeypatch.setattr(graph, "ENABLED", True)
    with torch.no_grad():
        got = [model(b["space_samples_src"], b["surface_samples_inputs"]) for b in batches]
    (e,) = graph._FWD_STATE[model].values()
    assert e.graph is not None and not e.failed
    assert got[4].data_ptr() != got[5].data_ptr()         # results are copies, not views of the graph's static output
    for g, w in zip(got, want):
        assert float((g.cpu() - w).norm(dim=-1).mean()) < 2e-6
    # gradients enabled, or train mode: never replayed
    out = model(batches[0]["space_samples_src"].requires_grad_(True), batches[0]["surface_samples_inputs"])
    assert out.requires_grad
