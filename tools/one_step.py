"""One training step at the bench configuration between cudaProfilerStart/Stop: the target of the ncu captures
(`ncu --profile-from-start off ...`). Same model / batch / step function as bench.py."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nsdp_b200 import synth
from nsdp_b200.model import build_model, optimizer_factory
dev = "cuda:0"
cfg = synth.make_config("forward")
model, train_on_batch, _, _ = build_model(cfg, device=dev)
schema = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
model.load_state_dict(synth.named_state_dict(schema, seed=0)); model.train()
_, opt = optimizer_factory(cfg["training"], model.parameters())
B = int(os.environ.get("B", 8))
batch = {k: v.to(dev) for k, v in synth.forward_batch(B, 4096, 50000, seed=1234).items()}
for _ in range(int(os.environ.get("WARM", 2))):
    train_on_batch(model, opt, batch, cfg)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = train_on_batch(model, opt, batch, cfg)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", loss)
