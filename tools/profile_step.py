"""torch.profiler table of one training step at the bench configuration (scratch tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from nsdp_b200 import synth
from nsdp_b200.model import build_model, optimizer_factory
dev = "cuda:0"
cfg = synth.make_config("forward")
model, train_on_batch, _, _ = build_model(cfg, device=dev)
schema = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
model.load_state_dict(synth.named_state_dict(schema, seed=0)); model.train()
_, opt = optimizer_factory(cfg["training"], model.parameters())
batch = {k: v.to(dev) for k, v in synth.forward_batch(8, 4096, 50000, seed=1).items()}
for _ in range(3): train_on_batch(model, opt, batch, cfg)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2): train_on_batch(model, opt, batch, cfg)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
