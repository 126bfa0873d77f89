import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from nsdp_b200 import ops
from test_gpu_vattn import _rand_case, vattn_reference, _rel_err
DEV = "cuda:0"
cases = [dict(B=2, M=333, N=100, K=7, D=200, has_global=True), dict(B=2, M=100, N=100, K=100, D=256, group_all=True),
         dict(B=2, M=150, N=150, K=10, D=120)]
for cfg in cases:
    case = _rand_case(seed=11, **cfg)
    names = [k for k, v in case.items() if torch.is_tensor(v) and v.is_floating_point()]
    cpu = {k: (v.double().clone().requires_grad_(True) if k in names else v) for k, v in case.items()}
    want = vattn_reference(sign=1.0, **cpu)
    go = torch.randn(want.shape, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    want.backward(go)
    res = {}
    for impl in (1, 0):
        ops.VATTN_IMPL = impl
        dev = {k: (v.to(DEV).contiguous().requires_grad_(True) if k in names else (v.to(DEV) if torch.is_tensor(v) else v))
               for k, v in case.items()}
        got = ops.vector_attention(sign=1.0, **dev)
        got.backward(go.float().to(DEV))
        res[impl] = {k: _rel_err(dev[k].grad, cpu[k].grad) for k in names}
        res[impl]["out"] = _rel_err(got.detach(), want.detach())
    print(cfg)
    for k in names + ["out"]:
        print(f"   {k:6s} ffma {res[1][k]:.2e}   tc {res[0][k]:.2e}")
