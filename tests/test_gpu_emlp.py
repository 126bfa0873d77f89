"""ElementwiseMLP (model/encoder/blocks.py:137-159) through nsdp_emlp_fwd_f32 / _bwd_f32 against
the CPU oracle's restatement (oracle.tdnet_oracle.elementwise_mlp, itself pinned to the live reference by
tests/test_oracle_golden.py): outputs, input gradient, every parameter gradient, running statistics and
num_batches_tracked, in training mode (batch statistics) and eval mode (running statistics), fp64 oracle as the truth."""
import numpy as np
import pytest
import torch

from nsdp_b200 import synth
from nsdp_b200.model.encoder.blocks import ElementwiseMLP
from oracle import tdnet_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("training", [True, False], ids=["train", "eval"])
@pytest.mark.parametrize("B,n,C", [(8, 500, 120), (8, 100, 256), (3, 37, 256), (2, 500, 64)])
def test_elementwise_mlp_forward_backward_against_oracle(B, n, C, training):
    from nsdp_b200 import ops
    mod = ElementwiseMLP(C)
    sd = synth.named_state_dict([("m." + k, tuple(v.shape)) for k, v in mod.state_dict().items()], seed=3)
    mod.load_state_dict({k[2:]: v for k, v in sd.items()})
    mod = mod.to(DEV).train(training)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, n, C, generator=g) * 1.5 + 0.3
    up = torch.randn(B, n, C, generator=g)
    # truth: the oracle in fp64 (autograd gives the gradients)
    sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    params = {k: v.requires_grad_(True) for k, v in sd64.items() if k.rsplit(".", 1)[-1] in ("weight", "bias")}
    x64 = x.double().requires_grad_(True)
    want = orc.elementwise_mlp(sd64, "m", x64, training=training)
    want.backward(up.double())
    # product
    before = ops.LAUNCHES
    xg = x.to(DEV).requires_grad_(True)
    got = mod(xg)
    got.backward(up.to(DEV))
    assert ops.LAUNCHES - before == 2                     # one fused forward op + one fused backward op, no torch glue
    assert _rel(got.detach().cpu().numpy(), want.detach().numpy()) < 2e-6
    assert _rel(xg.grad.cpu().numpy(), x64.grad.numpy()) < 2e-5
    for k, p in mod.named_parameters():
        ref = params["m." + k].grad.numpy()
        if k in ("conv1.bias", "conv2.bias") and training:
            # a bias feeding straight into a batch-statistics BatchNorm has a true gradient of 0: both sides are rounding noise
            assert np.abs(p.grad.cpu().numpy()).max() < 1e-4 * max(np.abs(up.numpy()).sum() / C, 1.0)
            continue
        assert _rel(p.grad.cpu().numpy(), ref) < 5e-5, k
    for k, v in mod.named_buffers():
        ref = sd64["m." + k]
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(ref) == (1 if training else 0)
        else:
            np.testing.assert_allclose(v.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


def test_elementwise_mlp_rejects_cpu_tensors():
    mod = ElementwiseMLP(64)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        mod(torch.randn(2, 10, 64))
