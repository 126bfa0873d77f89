"""Forward time of the fused MLP at W = 16 / 32 / 64 for the kernel geometry selected by NSDP_MLP_NARROW (tuning runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nsdp_b200 import ops, synth
DEV = "cuda:0"
R, L, NBUF = 1_000_000, 6, 12
g = torch.Generator().manual_seed(5)
xs = [(torch.rand(R, 3, generator=g) - 0.5).to(DEV) for _ in range(NBUF)]
outs = [torch.empty(R, 3, device=DEV) for _ in range(NBUF)]
res = []
for W in (16, 32, 64):
    w = [torch.from_numpy(t).to(DEV) for t in synth.mlp_weights(W, L, seed=W)]
    net = ops.FusedMLP(*w, impl=0)
    for i in range(5):
        net(xs[i % NBUF], outs[i % NBUF])
    ts = []
    # 24 calls queued behind a long-running dummy kernel per event pair: the host's launch cost stays off the clock
    big = torch.empty(64 << 20, device=DEV)
    for rep in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(4):
            big.add_(1.0)
        a.record()
        for i in range(24):
            net(xs[i % NBUF], outs[i % NBUF])
        b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) / 24)
    ts.sort()
    res.append(f"W={W}: {ts[len(ts) // 2] * 1e3:.1f} us")
print("variant", os.environ.get("NSDP_MLP_NARROW", "0"), "  ".join(res))
