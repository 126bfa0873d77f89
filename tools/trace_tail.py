"""Timeline of CTA 0 of the tail backward chain kernel (trace build)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
dev = "cuda:0"
trace = torch.zeros(16384, dtype=torch.int64, device=dev)
os.environ["NSDP_TRACE_PTR"] = str(trace.data_ptr())
from nsdp_b200 import ops
R, C, nb, O, H = 400000, 200, 5, 3, 128
g = torch.Generator().manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g)
args = [r(R, C), r(C, (1 + nb) * H) / np.sqrt(C), r((1 + nb) * H) * 0.1, r(nb, H, H) / np.sqrt(H), r(nb, H) * 0.1,
        r(nb, H, H) / np.sqrt(H), r(nb, H) * 0.1, r(H, O) / np.sqrt(H), r(O) * 0.1]
devs = [a.to(dev).requires_grad_(True) for a in args]
go = torch.randn(R, O, generator=g).to(dev)
for it in range(2):
    trace.zero_()
    out = ops.resnet_tail(*devs)
    out.backward(go)
torch.cuda.synchronize()
t = trace.cpu().tolist()
n = min(t[0], 4000)
ev = sorted([(t[2 + 2 * i], t[1 + 2 * i]) for i in range(n)])
t0 = ev[0][0]
names = {100: "M wait operand", 101: "M operand ready", 102: "M GEMM(s) issued", 200: "W wait acc", 201: "W acc ready", 202: "W published"}
last = {}
for c, e in ev[:150]:
    role = "M" if e < 200 else "W"
    d = c - last.get(role, c); last[role] = c
    print(f"{c - t0:9d} (+{d:7d}) {names.get(e, e)}")
