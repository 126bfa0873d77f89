"""Prints the gradient errors of one training step (tensor-core path) against the live-reference goldens
(tests/golden/, 'train_fwd_*'): used to judge precision switches such as NSDP_DW_TERMS."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nsdp_b200 import synth
from nsdp_b200.model import build_model
from nsdp_b200.model.utils import compute_l2_error
DEV = "cuda:0"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
golden = np.load(os.path.join(root, "tests/golden/tdnet_reference.npz"))
schemas = json.load(open(os.path.join(root, "tests/golden/state_dict_schema.json")))
model, *_ = build_model(synth.make_config("forward"), device=DEV)
model.load_state_dict(synth.named_state_dict([(k, s) for k, s in schemas["forward"]], seed=0))
model.train()
b = synth.forward_batch(2, 768, 640, seed=5, fp16_grid=False)
q = b["space_samples_src"].to(DEV).requires_grad_(True)
surf = b["surface_samples_inputs"].to(DEV).requires_grad_(True)
pred = model(q, surf)
loss = compute_l2_error(pred, b["space_samples_tgt"].to(DEV))
loss.backward()
rel = lambda a, ref: float(np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30))
print("terms", os.environ.get("NSDP_DW_TERMS", "3"), "dq", rel(q.grad.cpu().numpy(), golden["train_fwd_dq"]),
      "dsurf", rel(surf.grad.cpu().numpy(), golden["train_fwd_dsurf"]))
grads = {k: p.grad for k, p in model.named_parameters()}
rows = []
for key in golden.files:
    if key.startswith("train_fwd_grad::"):
        k = key.split("::", 1)[1]
        rows.append((rel(grads[k].cpu().numpy(), golden[key]), k))
rows.sort(reverse=True)
for r, k in rows[:12]:
    print(f"  {r:.3e}  {k}")
names = [k for k, _ in model.named_parameters()]
norms = np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in model.named_parameters()])
ref = golden["train_fwd_gradnorms"]
dev = [(abs(a - r) / max(r, 1e-30), n) for n, a, r in zip(names, norms, ref) if r > 1e-6]
dev.sort(reverse=True)
print("  worst grad-norm deviations:", [(f"{d:.2e}", n) for d, n in dev[:5]])
