"""Which cuBLAS GEMMs of a training step cost what: torch profiler with shapes over one EAGER step (NSDP_B200_GRAPH=0)."""
import os, sys
os.environ["NSDP_B200_GRAPH"] = "0"
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
for _ in range(3):
    train_on_batch(model, opt, dict(batch), cfg)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    train_on_batch(model, opt, dict(batch), cfg)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key in ("aten::mm", "aten::addmm", "aten::bmm", "aten::linear", "aten::matmul")]
rows.sort(key=lambda e: -e.device_time_total)
tot = 0.0
for e in rows[:24]:
    if e.key in ("aten::mm", "aten::addmm", "aten::bmm"):
        tot += e.device_time_total
        print(f"{e.device_time_total / 1e3:8.3f} ms  x{e.count:3d}  {e.key:12s} {e.input_shapes}")
print("sum of listed mm/addmm/bmm:", round(tot / 1e3, 3), "ms")
print("---- other ATen ops by device time (self)")
others = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith("aten::") and e.key not in ("aten::mm", "aten::addmm", "aten::bmm") and e.self_device_time_total > 0]
others.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in others)
print("total", round(tot / 1e3, 3), "ms over", sum(e.count for e in others), "calls")
for e in others[:28]:
    print(f"{e.self_device_time_total / 1e3:8.3f} ms  x{e.count:3d}  {e.key:28s} {str(e.input_shapes)[:110]}")
