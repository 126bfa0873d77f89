"""CUDA-event timing of the decoder forward (eval) and forward + backward (train) at the bench shape; for A/B runs of
library variants on the SAME box: NSDP_B200_LIB=<variant .so> python tools/time_decode.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = "cuda:0"
from nsdp_b200 import synth
from nsdp_b200.model import build_model
B, N, Q = 8, 4096, 50000
model, *_ = build_model(synth.make_config("forward"), device=dev)
schema = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
model.load_state_dict(synth.named_state_dict(schema, seed=0))
batch = {k: v.to(dev) for k, v in synth.forward_batch(B, N, Q, seed=1).items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


model.eval()
with torch.no_grad():
    enc = model.encode(batch["surface_samples_inputs"])
    t_fwd = timed(lambda: model.decode(batch["space_samples_src"], enc))
model.train()
enc = {k: v.detach().requires_grad_(k != "anchors") for k, v in enc.items()}


def fb():
    out = model.decode(batch["space_samples_src"], enc)
    out.square().mean().backward()


t_fb = timed(fb)
print(f"{os.environ.get('NSDP_B200_LIB', 'default')}: decoder forward (eval) {t_fwd:.3f} ms, forward + backward {t_fb:.3f} ms")
if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            fb()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]
    for e in rows:
        print(f"  {e.device_time_total / 3e3:8.3f} ms/iter  {e.count // 3:3d} launches  {e.key[:90]}")
