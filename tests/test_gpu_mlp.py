"""Parity of the fused neural-field MLP (BASELINE.json configs[3], nsdp_fused_mlp_fwd_f32) through the C ABI: the tcgen05
kernel and the fp32 CUDA-core kernel against the fp64 numpy oracle (oracle/mlp_oracle.py), ragged row counts, every
instantiated width, and the full 1M x 256 x 8-layer size through a row-sampled check (rows are independent)."""
import numpy as np
import pytest
import torch

from nsdp_b200 import ops, synth
from oracle import mlp_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net(W, L, impl, Cin=3, O=3, seed=0):
    w = synth.mlp_weights(W, L, Cin=Cin, O=O, seed=seed)
    net = ops.FusedMLP(*[torch.from_numpy(t).to(DEV) for t in w], impl=impl)
    return w, net


def _rel(y, ref):
    """Error relative to the part of the output that VARIES over rows (the mean is mostly the last bias)."""
    den = np.linalg.norm(ref - ref.mean(0)) if len(ref) > 16 else 0.3 * np.linalg.norm(ref)
    return float(np.linalg.norm(y.astype(np.float64) - ref) / den)


@pytest.mark.parametrize("W", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("R", [1, 127, 129, 20011])
def test_tcgen05_kernel_matches_oracle(W, R):
    w, net = _net(W, 6, impl=2, seed=W)
    x = (np.random.default_rng(R).random((R, 3), dtype=np.float32) - 0.5)
    y = net(torch.from_numpy(x).to(DEV))
    y2 = net(torch.from_numpy(x).to(DEV))          # second call re-uses the packed weight image
    torch.cuda.synchronize()
    ref = mlp_oracle.mlp_forward(x, *w)
    assert _rel(y.cpu().numpy(), ref) < 5e-5       # bf16x3 split precision: fp32-grade (plain bf16 is ~1e-2 here)
    assert torch.equal(y, y2)


@pytest.mark.parametrize("W,L,Cin,O", [(16, 1, 3, 3), (64, 7, 4, 4), (128, 2, 1, 1), (256, 3, 2, 2)])
def test_tcgen05_kernel_shapes(W, L, Cin, O):
    w, net = _net(W, L, impl=2, Cin=Cin, O=O, seed=7)
    x = (np.random.default_rng(1).random((3001, Cin), dtype=np.float32) - 0.5)
    y = net(torch.from_numpy(x).to(DEV)).cpu().numpy()
    assert _rel(y, mlp_oracle.mlp_forward(x, *w)) < 1e-4


@pytest.mark.parametrize("W,L", [(16, 6), (100, 2), (256, 6), (64, 0)])
def test_fp32_kernel_matches_oracle(W, L):
    w, net = _net(W, L, impl=1, seed=3)
    x = (np.random.default_rng(2).random((1037, 3), dtype=np.float32) - 0.5)
    y = net(torch.from_numpy(x).to(DEV)).cpu().numpy()
    assert _rel(y, mlp_oracle.mlp_forward(x, *w)) < 2e-5


def test_unsupported_width_is_refused_by_the_tcgen05_path():
    w, net = _net(100, 2, impl=2, seed=3)
    with pytest.raises(RuntimeError, match="unsupported"):
        net(torch.zeros(8, 3, device=DEV))


def test_c4_full_size_rows_sampled():
    """1 000 000 query points x width 256 x 8 linear layers (configs[3]); the oracle checks 4096 sampled rows."""
    R, W, L = 1_000_000, 256, 6
    w, net = _net(W, L, impl=0, seed=11)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(R, 3, generator=g) - 0.5
    y = net(x.to(DEV)).cpu().numpy()
    rows = np.random.default_rng(0).choice(R, 4096, replace=False)
    rows[:2] = (0, R - 1)
    ref = mlp_oracle.mlp_forward(x.numpy()[rows], *w)
    assert _rel(y[rows], ref) < 5e-5
    assert np.isfinite(y).all()


def _torch_reference_grads(x, w, d_out):
    """fp64 autograd of the nn.Linear / ReLU stack on the CPU (what the reference's decoder pattern does under autograd)."""
    w_in, b_in, w_h, b_h, w_out, b_out = [torch.from_numpy(np.asarray(t, np.float64)).requires_grad_(True) for t in w]
    xt = torch.from_numpy(x.astype(np.float64)).requires_grad_(True)
    h = torch.relu(xt @ w_in.t() + b_in)
    for l in range(w_h.shape[0]):
        h = torch.relu(h @ w_h[l].t() + b_h[l])
    out = h @ w_out.t() + b_out
    out.backward(torch.from_numpy(d_out.astype(np.float64)))
    return [t.grad.numpy() for t in (xt, w_in, b_in, w_h, b_h, w_out, b_out)]


def _run_backward(W, L, R, Cin=3, O=3, seed=0):
    w = synth.mlp_weights(W, L, Cin=Cin, O=O, seed=seed)
    rng = np.random.default_rng(seed + R)
    x = rng.random((R, Cin), dtype=np.float32) - 0.5
    d_out = rng.standard_normal((R, O)).astype(np.float32)
    # ReLU gradients are discontinuous at z = 0: rows with a pre-activation within 1e-4 of zero (where fp32-grade and fp64
    # arithmetic may disagree about the sign) get a zero upstream gradient on BOTH sides, so they test nothing and break nothing
    h = x.astype(np.float64) @ w[0].astype(np.float64).T + w[1]
    zmin = np.abs(h).min(1)
    for l in range(L):
        h = np.maximum(h, 0) @ w[2][l].astype(np.float64).T + w[3][l]
        zmin = np.minimum(zmin, np.abs(h).min(1))
    d_out[zmin < 1e-4] = 0.0
    assert (zmin >= 1e-4).mean() > 0.5
    w_in, b_in, w_h, b_h, w_out, b_out = [torch.from_numpy(t).to(DEV) for t in w]
    params = [w_in.t().contiguous(), b_in, w_h.transpose(1, 2).contiguous(), b_h, w_out.t().contiguous(), b_out]
    params = [p.requires_grad_(True) for p in params]
    xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
    out = ops.fused_mlp(xt, *params)
    out.backward(torch.from_numpy(d_out).to(DEV))
    torch.cuda.synchronize()
    got = [xt.grad.cpu().numpy(), params[0].grad.t().cpu().numpy(), params[1].grad.cpu().numpy(),
           params[2].grad.transpose(1, 2).cpu().numpy(), params[3].grad.cpu().numpy(), params[4].grad.t().cpu().numpy(),
           params[5].grad.cpu().numpy()]
    return got, _torch_reference_grads(x, w, d_out), out.detach().cpu().numpy(), mlp_oracle.mlp_forward(x, *w)


@pytest.mark.parametrize("W,L,R", [(16, 2, 1000), (32, 6, 2049), (64, 6, 5000), (128, 3, 4097), (256, 6, 3000), (256, 1, 129)])
def test_backward_matches_fp64_autograd(W, L, R):
    got, want, out, out_ref = _run_backward(W, L, R, seed=W + L)
    assert _rel(out, out_ref) < 5e-5
    names = ["d_x", "d_w_in", "d_b_in", "d_w_h", "d_b_h", "d_w_out", "d_b_out"]
    rels = {}
    for n, g, r in zip(names, got, want):
        rels[n] = float(np.linalg.norm(g.astype(np.float64) - r) / max(np.linalg.norm(r), 1e-30))
    print(f"W={W} L={L} R={R}: " + " ".join(f"{n} {v:.1e}" for n, v in rels.items()))
    assert max(rels.values()) < 1e-4, rels        # measured ~1e-5 (SURVEY 8d asks for < 1e-3)


def test_backward_multi_segment_and_narrow_io():
    """More tiles than one staging segment (NSDP_MLP_SEG tiles) and Cin = O = 1."""
    import os
    assert int(os.environ.get("NSDP_MLP_SEG", "592")) * 128 < 100_000
    got, want, _, _ = _run_backward(64, 2, 100_003, Cin=1, O=1, seed=5)
    for g, r in zip(got, want):
        rel = float(np.linalg.norm(g.astype(np.float64) - r) / max(np.linalg.norm(r), 1e-30))
        assert rel < 1e-4, rel
