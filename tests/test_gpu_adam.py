"""nsdp_b200.optim.Adam (csrc/adam.cu) against torch.optim.Adam: same trajectory, same state_dict layout, capturable."""
import copy

import pytest
import torch

from nsdp_b200.optim import Adam

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SHAPES = [(256, 256), (1,), (120,), (3, 200), (4097,), (65537,), (208, 128), (7, 5, 3)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in SHAPES]


def _grads(params, seed):
    g = torch.Generator().manual_seed(seed)
    for i, p in enumerate(params):
        p.grad = None if i == 2 and seed % 2 else (torch.randn(*p.shape, generator=g) * (10.0 ** (i % 3 - 2))).to(DEV)


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_adam_matches_torch(wd):
    ours, ref = _params(0), _params(0)
    o1 = Adam([{"params": ours, "lr": 5e-4, "weight_decay": wd}])
    o2 = torch.optim.Adam([{"params": ref, "lr": 5e-4, "weight_decay": wd}])
    for it in range(12):
        _grads(ours, it); _grads(ref, it)
        if it == 6:
            for o in (o1, o2):
                o.param_groups[0]["lr"] = 5e-5                       # model/learningrate.py adjust_learning_rate
        o1.step(); o2.step()
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), float((a - b).abs().max())
    s1, s2 = o1.state_dict(), o2.state_dict()
    assert s1["state"].keys() == s2["state"].keys()
    for k in s1["state"]:
        assert set(s1["state"][k]) == set(s2["state"][k]) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(s1["state"][k]["step"]) == float(s2["state"][k]["step"])
        assert torch.allclose(s1["state"][k]["exp_avg_sq"], s2["state"][k]["exp_avg_sq"], rtol=2e-6, atol=1e-12)
    # checkpoints travel both ways: the default torch optimizer's state into ours, ours into torch's capturable fused Adam,
    # then both continue identically
    o3 = torch.optim.Adam([{"params": ref, "lr": 5e-5, "weight_decay": wd}], fused=True, capturable=True)
    o3.load_state_dict(copy.deepcopy(s1))
    o1.load_state_dict(copy.deepcopy(s2))
    for a, b in zip(ours, ref):
        b.data.copy_(a.data)
    for it in range(12, 15):
        _grads(ours, it); _grads(ref, it)
        o1.step(); o3.step()
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7)
    assert float(o1.state[ours[0]]["step"]) == float(o3.state[ref[0]]["step"]) == 15.0


def test_adam_step_is_capturable():
    ours, ref = _params(1), _params(1)
    o1 = Adam([{"params": ours, "lr": 1e-3}])
    o2 = torch.optim.Adam([{"params": ref, "lr": 1e-3}])
    _grads(ours, 0); _grads(ref, 0)
    for p in ours + ref:
        assert p.grad is not None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        o1.step()
    torch.cuda.current_stream().wait_stream(side)
    o2.step()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        o1.step()
    for it in range(4):
        gen = torch.Generator().manual_seed(100 + it)
        for a, b in zip(ours, ref):
            new = torch.randn(*a.shape, generator=gen).to(DEV)
            a.grad.copy_(new); b.grad.copy_(new)
        g.replay(); o2.step()
    torch.cuda.synchronize()
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7)
    assert float(o1.state[ours[0]]["step"]) == 5.0


@pytest.mark.parametrize("R,K,N,bias", [(32768, 4, 360, True), (5000, 3, 120, False), (4096, 8, 129, True), (70000, 1, 7, True)])
def test_linear_narrow_backward(R, K, N, bias):
    """ops.linear (narrow input, many rows): the weight / bias gradient kernel against torch autograd in float64."""
    from nsdp_b200 import ops
    g = torch.Generator().manual_seed(R + K)
    x = torch.randn(R, K, generator=g).to(DEV).requires_grad_(True)
    w = torch.randn(N, K, generator=g).to(DEV).requires_grad_(True)
    b = torch.randn(N, generator=g).to(DEV).requires_grad_(True) if bias else None
    d_y = torch.randn(R, N, generator=g).to(DEV)
    y = ops.linear(x, w, b)
    y.backward(d_y)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    yd = torch.nn.functional.linear(xd, wd, bd)
    yd.backward(d_y.double())
    assert torch.allclose(y.double(), yd, rtol=1e-5, atol=1e-5)
    rel = lambda a, t: float((a.double() - t).norm() / t.norm())
    assert rel(w.grad, wd.grad) < 2e-6 and rel(x.grad, xd.grad) < 2e-6
    if bias:
        assert rel(b.grad, bd.grad) < 2e-6
