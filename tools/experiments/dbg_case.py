import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from nsdp_b200 import ops
from test_gpu_vattn import _rand_case, _rel_err, vattn_reference
DEV = "cuda:0"
def run(cfg, impl):
    ops.VATTN_IMPL = impl
    case = _rand_case(seed=11, **cfg)
    names = [k for k, v in case.items() if torch.is_tensor(v) and v.is_floating_point()]
    cpu = {k: (v.double().clone().requires_grad_(True) if k in names else v) for k, v in case.items()}
    dev = {k: (v.to(DEV).contiguous().requires_grad_(True) if k in names else (v.to(DEV) if torch.is_tensor(v) else v)) for k, v in case.items()}
    want = vattn_reference(sign=1.0, **cpu); got = ops.vector_attention(sign=1.0, **dev)
    go = torch.randn(want.shape, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    want.backward(go); got.backward(go.float().to(DEV))
    print(cfg, "impl", impl, "fwd", f"{_rel_err(got.detach(), want.detach()):.1e}", {k: f"{_rel_err(dev[k].grad, cpu[k].grad):.1e}" for k in names})
base = dict(B=2, M=50, N=111, K=5, D=160, has_global=True, shape_query=True)
for over in [{}, {"D": 200}, {"K": 7}, {"N": 100}, {"M": 64}, {"shape_query": False}, {"has_global": False, "shape_query": False}]:
    for impl in (1, 0):
        run({**base, **over}, impl)
