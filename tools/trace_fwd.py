"""Timeline of CTA 0 of the decoder forward chain kernel (trace build: NSDP_BUILD_VARIANT=trace, -DNSDP_TRACE)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = "cuda:0"
trace = torch.zeros(16384, dtype=torch.int64, device=dev)
os.environ["NSDP_TRACE_FWD_PTR"] = str(trace.data_ptr())
from nsdp_b200 import synth
from nsdp_b200.model import build_model
B, N, Q = 8, 4096, 50000
model, *_ = build_model(synth.make_config("forward"), device=dev)
schema = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
model.load_state_dict(synth.named_state_dict(schema, seed=0)); model.eval()
batch = {k: v.to(dev) for k, v in synth.forward_batch(B, N, Q, seed=1).items()}
with torch.no_grad():
    enc = model.encode(batch["surface_samples_inputs"])
    for it in range(2):
        trace.zero_()
        out = model.decode(batch["space_samples_src"], enc)
torch.cuda.synchronize()
t = trace.cpu().tolist()
n = min(t[0], 4000)
ev = [(t[1 + 2 * i], t[2 + 2 * i]) for i in range(n)]
ev.sort(key=lambda e: e[1])
t0 = ev[0][1]
names = {100: "M wait acc1 free (GEMM1b)", 101: "M acc1 free", 102: "M GEMM1b issued, wait acc0 free", 103: "M acc0 free (GEMM1a)",
         110: "M GEMM1a issued, wait acc0 free (GEMM2)", 111: "M acc0 free", 113: "M GEMM2 issued",
         200: "W tile start", 202: "W GEMM1 done seen", 203: "W acc0 in registers", 204: "W G written",
         205: "W next row info loaded", 206: "W GEMM2 done seen", 208: "W s in registers + next operands written", 207: "W epilogue 2 done"}
last = {}
print("CTA 0; cycles since start, delta to previous event of the same role")
for i, (e, c) in enumerate(ev[:120]):
    role = "M" if e < 200 else "W"
    d = c - last.get(role, c)
    last[role] = c
    print(f"{c - t0:9d} (+{d:7d}) {names.get(e, e)}")
