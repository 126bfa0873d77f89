"""N>1 host logic on CPU: world_size-2 gloo. Checks that the flat gradient all-reduce averages across ranks
(including parameters that received no gradient on some rank), that replicas are synchronised by
broadcast_parameters, and the contiguous batch sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from nsdp_b200 import dist as nd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                      # replicas start DIFFERENT ...
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2, bias=False), torch.nn.Linear(2, 2))
    nd.broadcast_parameters(model)                     # ... and are made identical
    w0 = model[0].weight.detach().clone()
    data = {"x": torch.arange(8 * 4, dtype=torch.float32).reshape(8, 4), "tag": "keep"}
    shard = nd.shard_batch(data)
    # the last layer is unused -> its grads stay None on every rank (like the pos_only block's q/k/v weights)
    loss = model[1](model[0](shard["x"])).pow(2).mean()
    loss.backward()
    local = model[0].weight.grad.clone()
    nd.allreduce_gradients(model)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    td.all_gather(gathered, local)
    q.put((rank, w0, shard["x"][:, 0].tolist(), shard["tag"], model[0].weight.grad.clone(), sum(gathered) / world,
           model[2].weight.grad.clone()))
    td.destroy_process_group()


def test_dp_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, w0a, x0, tag0, g0, mean0, unused0), (r1, w0b, x1, tag1, g1, mean1, unused1) = res
    assert torch.equal(w0a, w0b)                                   # broadcast made replicas identical
    assert x0 == [0.0, 4.0, 8.0, 12.0] and x1 == [16.0, 20.0, 24.0, 28.0] and tag0 == "keep"
    torch.testing.assert_close(g0, mean0)
    torch.testing.assert_close(g0, g1)                             # every rank holds the averaged gradient
    assert torch.count_nonzero(unused0) == 0 and torch.count_nonzero(unused1) == 0


def _decode_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    pts = torch.arange(2 * 7 * 3, dtype=torch.float32).reshape(2, 7, 3)    # Q = 7 is not divisible by 2: padded slice
    seen = []

    def decode(p):                                                          # stand-in for model.decode(p, encoding)
        seen.append(p.shape[1])
        return p * 2.0 + 1.0

    out = nd.sharded_decode(decode, pts)
    q.put((rank, out, seen))
    td.destroy_process_group()


def test_query_sharded_decode_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_decode_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.arange(2 * 7 * 3, dtype=torch.float32).reshape(2, 7, 3) * 2.0 + 1.0
    for rank, out, seen in res:
        assert torch.equal(out, want)          # every rank ends with all queries, in order
        assert seen == [4]                     # ... having decoded only its (padded) slice


def _mlp():
    torch.manual_seed(11)
    return torch.nn.Sequential(torch.nn.Linear(4, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                               torch.nn.Linear(16, 3), torch.nn.Linear(3, 3))     # the last layer is never used


def _step_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      NSDP_B200_BUCKET_BYTES="256")          # tiny buckets: several collectives leave DURING backward
    td.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(50 + rank)                              # replicas start different, as after a botched resume
    model = torch.nn.Sequential(torch.nn.Linear(4, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                                torch.nn.Linear(16, 3), torch.nn.Linear(3, 3))
    if rank == 0:
        model.load_state_dict(_mlp().state_dict())
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    x = torch.randn(8, 4, generator=torch.Generator().manual_seed(1))
    y = torch.randn(8, 3, generator=torch.Generator().manual_seed(2))
    mine = nd.shard_batch({"x": x, "y": y})
    launched_early = []
    for step in range(3):
        nd.zero_grad(model, opt)                               # first call: rank 0's weights/optimizer state become everybody's
        bk = nd._buckets_for(model)
        loss = (model[:5](mine["x"]) - mine["y"]).pow(2).mean()
        loss.backward()
        launched_early.append(bk.launched)                     # buckets already reduced/in flight when backward returns
        nd.allreduce_gradients(model)
        opt.step()
    q.put((rank, {k: v.detach().numpy() for k, v in model.state_dict().items()}, launched_early, bk.nb,
           all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bk.params, bk.views))))
    td.destroy_process_group()


def test_bucketed_overlapped_allreduce_matches_single_process_training_world2():
    """3 Adam steps on 2 ranks x 4 rows == 3 steps of one process on the 8 rows; gradients live in the flat buffer; from the
    second step on buckets leave during backward."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_step_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _mlp()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-2)
    x = torch.randn(8, 4, generator=torch.Generator().manual_seed(1))
    y = torch.randn(8, 3, generator=torch.Generator().manual_seed(2))
    for step in range(3):
        opt.zero_grad()
        (ref[:5](x) - y).pow(2).mean().backward()
        opt.step()
    for rank, sd, early, nb, views_ok in res:
        assert views_ok and nb >= 3
        assert early[0] == 0                       # step 1 learns which parameters get gradients: everything at the end
        assert early[1] >= 1 and early[2] >= 1      # afterwards collectives leave while backward is still running
        for k, v in ref.state_dict().items():
            torch.testing.assert_close(torch.from_numpy(sd[k]), v, atol=1e-6, rtol=1e-5)


def _net():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(5, 6), torch.nn.BatchNorm1d(6), torch.nn.ReLU(), torch.nn.Linear(6, 2))


def _syncbn_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    model = nd.convert_sync_batchnorm(_net())
    model.train()
    x = torch.randn(12, 5, generator=torch.Generator().manual_seed(9))
    mine = nd.shard_batch({"x": x})["x"]
    out = model(mine)
    out.pow(2).mean().backward()
    nd.allreduce_gradients(model)
    # numpy payloads are pickled by value (torch tensors travel as shared-memory handles that die with the worker)
    q.put((rank, out.detach().numpy(), {k: v.grad.numpy() for k, v in model.named_parameters()},
           {k: v.numpy() for k, v in model.named_buffers()}, list(model.state_dict().keys())))
    td.destroy_process_group()


def test_syncbn_reproduces_the_single_process_batch_world2():
    """SURVEY.md §8e optional syncbn mode: 2 ranks x 6 rows == 1 process x 12 rows (outputs, gradients, running stats)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_syncbn_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _net()
    ref.train()
    x = torch.randn(12, 5, generator=torch.Generator().manual_seed(9))
    out = ref(x)
    out.pow(2).mean().backward()
    torch.testing.assert_close(torch.from_numpy(np.concatenate([res[0][1], res[1][1]])), out.detach(), atol=1e-5, rtol=1e-5)
    for rank, _, grads, bufs, keys in res:
        assert keys == list(ref.state_dict().keys())
        for k, v in ref.named_parameters():
            torch.testing.assert_close(torch.from_numpy(grads[k]), v.grad, atol=1e-5, rtol=1e-4)
        for k, v in ref.named_buffers():
            torch.testing.assert_close(torch.from_numpy(bufs[k]), v, atol=1e-6, rtol=1e-5)


def test_inactive_without_process_group():
    model = torch.nn.Linear(2, 2)
    model(torch.ones(1, 2)).sum().backward()
    g = model.weight.grad.clone()
    nd.allreduce_gradients(model)                                  # no-op
    assert torch.equal(g, model.weight.grad)
    assert nd.shard_batch({"x": torch.zeros(4, 1)})["x"].shape[0] == 4
    with pytest.raises(ValueError):
        nd.shard_batch({"x": torch.zeros(5, 1)}, rank=0, world=2)
    pts = torch.rand(1, 5, 3)
    assert torch.equal(nd.sharded_decode(lambda p: p + 1.0, pts), pts + 1.0)   # no process group: plain decode


# ---------------------------------------------------------------------------------------------------------------------
# the REAL TDNet mirror through the real entry points (build_model joins the job, optimizer_factory, train_on_batch), with the
# header's kernel contracts standing in for the CUDA kernels (tests/helpers/contracts.py): SURVEY.md section 4 "distributed"
# ---------------------------------------------------------------------------------------------------------------------
def _tdnet_batch():
    from nsdp_b200 import synth
    b = synth.forward_batch(2, 384, 160, seed=41, fp16_grid=False)
    return {k: b[k].double() for k in ("surface_samples_inputs", "space_samples_src", "space_samples_tgt")}


def _tdnet_train(world_rank=None):
    """Two Adam steps of the forward TDNet in float64 on this process's share of the 2-shape batch; returns losses + state."""
    from helpers import contracts
    from nsdp_b200 import synth
    from nsdp_b200.model import build_model, optimizer_factory
    contracts.install()
    cfg = synth.make_config("forward")
    torch.manual_seed(7 if not world_rank else 1000 + world_rank)       # ranks > 0 start from OTHER weights: rank 0's must win
    model, train_on_batch, *_ = build_model(cfg, device="cpu")
    if not world_rank:
        schema = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
        model.load_state_dict(synth.named_state_dict(schema, seed=0))
        if world_rank == 0:
            nd.broadcast_parameters(model)                               # what train.py's checkpoint loading + first step amount to
    else:
        nd.broadcast_parameters(model)
    model.double()
    model.train()
    _, opt = optimizer_factory(cfg["training"], model.parameters())
    assert type(opt) is torch.optim.Adam                                # the library's fused Adam is for CUDA parameters only
    data = _tdnet_batch()
    if world_rank is not None:
        data = nd.shard_batch(data)
    losses = [train_on_batch(model, opt, data, cfg) for _ in range(2)]
    return losses, {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}


def _tdnet_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      NSDP_B200_SYNCBN="1", NSDP_B200_BUCKET_BYTES=str(1 << 20))
    torch.set_num_threads(2)
    losses, sd = _tdnet_train(rank)                                     # build_model creates the gloo group from the environment
    assert nd.is_active()
    q.put((rank, losses, sd))
    td.destroy_process_group()


def test_tdnet_two_ranks_with_syncbn_equal_one_process_on_the_whole_batch():
    """2 ranks x 1 shape (BatchNorm statistics over both ranks, averaged gradients, bucketed all-reduce from autograd hooks)
    == 1 process x 2 shapes, after two Adam steps of the real model: identical weights and BatchNorm buffers (float64: 1e-9),
    and the mean of the per-rank losses is the single-process loss."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tdnet_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref_losses, ref_sd = _tdnet_train(None)
    for step in range(2):
        assert abs(0.5 * (res[0][1][step] + res[1][1][step]) - ref_losses[step]) < 1e-9
    assert ref_losses[1] != ref_losses[0]                               # the optimizer moved the weights
    for rank, _, sd in res:
        assert list(sd.keys()) == list(ref_sd.keys())
        for k, v in ref_sd.items():
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(v) == 2, k
            else:
                np.testing.assert_allclose(sd[k], v, atol=1e-9, rtol=1e-7, err_msg=f"rank {rank}: {k}")
    for k in ref_sd:                                                    # and the replicas agree with each other bit for bit
        assert np.array_equal(res[0][2][k], res[1][2][k]), k
