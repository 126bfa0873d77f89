"""One configs[3] call (1M rows, W from argv, default 256) after two warm-up calls: the command the ncu captures wrap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nsdp_b200 import ops, synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = ops.FusedMLP(*[torch.from_numpy(t).to("cuda:0") for t in synth.mlp_weights(W, 6, seed=W)])
x = torch.rand(1_000_000, 3, device="cuda:0") - 0.5
y = net(x)
torch.cuda.synchronize(); torch.cuda.profiler.start()
for _ in range(int(os.environ.get("REPS", "3"))):
    y = net(x)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print(float(y.abs().mean()))
