"""Timeline of CTA 0 of the decoder backward chain kernel (trace build: NSDP_BUILD_VARIANT=trace, -DNSDP_TRACE)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = "cuda:0"
trace = torch.zeros(16384, dtype=torch.int64, device=dev)
os.environ["NSDP_TRACE_BWD_PTR"] = str(trace.data_ptr())
from nsdp_b200 import synth
from nsdp_b200.model import build_model
B, N, Q = 8, 4096, 50000
model, *_ = build_model(synth.make_config("forward"), device=dev)
schema = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
model.load_state_dict(synth.named_state_dict(schema, seed=0)); model.train()
batch = {k: v.to(dev) for k, v in synth.forward_batch(B, N, Q, seed=1).items()}
with torch.no_grad():
    enc = model.encode(batch["surface_samples_inputs"])
enc = {k: v.detach().requires_grad_(k != "anchors") for k, v in enc.items()}
for it in range(2):
    trace.zero_()
    out = model.decode(batch["space_samples_src"], enc)
    out.square().mean().backward()
torch.cuda.synchronize()
t = trace.cpu().tolist()
n = min(t[0], 4000)
ev = [(t[1 + 2 * i], t[2 + 2 * i]) for i in range(n)]
ev.sort(key=lambda e: e[1])
t0 = ev[0][1]
names = {200: "W tile start", 201: "W E/H written", 202: "W GEMM1 done seen", 203: "W G written", 204: "W GEMM2 done seen",
         205: "W da/ds written", 206: "W GEMM3 done seen", 207: "W ds operand written", 208: "W dgp epilogue done",
         209: "W GEMM4a done seen", 210: "W dgp operand written", 211: "W GEMM4b done seen", 212: "W dpre written (bar)",
         213: "W d_xyz done (bar)"}
for g in range(5):
    names[100 + 10 * g] = f"M wait acc free {g}"; names[101 + 10 * g] = f"M acc free {g}"; names[102 + 10 * g] = f"M GEMM {g} issued"
last = {}
print("first launch segment, CTA 0; cycles since start, delta to previous event of the same role")
for i, (e, c) in enumerate(ev[:140]):
    role = "M" if e < 200 else "W"
    d = c - last.get(role, c)
    last[role] = c
    print(f"{c - t0:9d} (+{d:7d}) {names.get(e, e)}")
