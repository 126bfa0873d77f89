"""Decoder cross-attention + tail forward at C2 size; target for ncu captures of the forward kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nsdp_b200 import synth
from nsdp_b200.model import build_model
dev = "cuda:0"
B, N, Q = int(os.environ.get("B", 8)), 4096, int(os.environ.get("Q", 50000))
model, *_ = build_model(synth.make_config("forward"), device=dev)
schema = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
model.load_state_dict(synth.named_state_dict(schema, seed=0)); model.eval()
batch = {k: v.to(dev) for k, v in synth.forward_batch(B, N, Q, seed=1).items()}
with torch.no_grad():
    enc = model.encode(batch["surface_samples_inputs"])
    reps = int(os.environ.get("REPS", 2))
    for it in range(reps):
        if it == reps - 1:
            torch.cuda.synchronize(); torch.cuda.profiler.start()
        out = model.decode(batch["space_samples_src"], enc)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ok")
