"""Gradients of the tensor-core kernels in the bf16-staged regime (million-row reductions) against the fp32 CUDA-core
kernels on the same inputs; prints relative L2 errors per gradient."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from nsdp_b200 import ops
from test_gpu_vattn import _rand_case, _rel_err
DEV = "cuda:0"

def run_vattn(impl, case, go):
    ops.VATTN_IMPL = impl
    names = [k for k, v in case.items() if torch.is_tensor(v) and v.is_floating_point()]
    dev = {k: (v.to(DEV).contiguous().requires_grad_(True) if k in names else (v.to(DEV) if torch.is_tensor(v) else v))
           for k, v in case.items()}
    out = ops.vector_attention(sign=1.0, **dev)
    out.backward(go)
    return out.detach(), {k: dev[k].grad for k in names}

case = _rand_case(B=2, M=70000, N=100, K=7, D=200, has_global=True, seed=3)
go = torch.randn(2, 70000, 200, generator=torch.Generator().manual_seed(5)).to(DEV)
o1, g1 = run_vattn(1, case, go)
o0, g0 = run_vattn(0, case, go)
print("vattn out err", _rel_err(o0, o1))
for k in g1:
    print(f"  vattn d_{k}: {_rel_err(g0[k], g1[k]):.3e}")

def run_tail(impl, args, go):
    ops.TAIL_IMPL = impl
    dev = [a.to(DEV).contiguous().requires_grad_(True) for a in args]
    out = ops.resnet_tail(*dev)
    out.backward(go)
    return out.detach(), [d.grad for d in dev]

R, C, nb, O, H = 300000, 200, 5, 3, 128
g = torch.Generator().manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g)
args = [r(R, C), r(C, (1 + nb) * H) / np.sqrt(C), r((1 + nb) * H) * 0.1, r(nb, H, H) / np.sqrt(H), r(nb, H) * 0.1,
        r(nb, H, H) / np.sqrt(H), r(nb, H) * 0.1, r(H, O) / np.sqrt(H), r(O) * 0.1]
go = torch.randn(R, O, generator=g).to(DEV)
o1, t1 = run_tail(1, args, go)
o0, t0 = run_tail(0, args, go)
print("tail out err", _rel_err(o0, o1))
for i, n in enumerate(["lat", "wc_t", "bc", "w0_t", "b0", "w1_t", "b1", "wo_t", "bo"]):
    print(f"  tail d_{n}: {_rel_err(t0[i], t1[i]):.3e}")
