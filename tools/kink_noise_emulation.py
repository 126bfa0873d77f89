"""Does a 3e-6 relative perturbation of every ReLU pre-activation (the bf16x3 product error) explain the tcgen05 path's
weight-gradient distance from the fp64 truth? fp32 oracle + injected noise vs tests/golden/tdnet_reference_r2.npz."""
import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import torch.nn.functional as F
from nsdp_b200 import synth
from oracle import tdnet_oracle as orc
from helpers_r2 import projection_vectors, masked_l2
torch.set_num_threads(8)
g = np.load('/root/repo/tests/golden/tdnet_reference_r2.npz')
schemas = json.load(open('/root/repo/tests/golden/state_dict_schema.json'))
NOISE = float(sys.argv[1]) if len(sys.argv) > 1 else 3e-6
gen = torch.Generator().manual_seed(0)
orig_mlp2 = orc._mlp2
def noisy_mlp2(sd, p, x):
    pre = orc._lin(sd, p + ".0", x)
    if NOISE > 0:
        pre = pre + pre.detach().pow(2).mean().sqrt() * NOISE * torch.randn(pre.shape, generator=gen)
    return orc._lin(sd, p + ".2", F.relu(pre))
orc._mlp2 = noisy_mlp2
cfg = synth.make_config("forward")["model"]
b = synth.forward_batch(2, 768, 640, seed=5, fp16_grid=False)
sd = synth.named_state_dict([(k, s) for k, s in schemas["forward"]], seed=0)
params = {k: v.requires_grad_(True) for k, v in sd.items() if k.rsplit(".", 1)[-1] in ("weight", "bias")}
pred = orc.tdnet_forward(sd, "", b["space_samples_src"], b["surface_samples_inputs"], cfg, False, training=True)
loss = masked_l2(pred, b["space_samples_tgt"], torch.from_numpy(g["fw64_keep"]))
loss.backward()
names = [str(n) for n in g["fw64_names"]]
rows = []
for i, n in enumerate(names):
    rn = float(g["fw64_gradnorms"][i])
    if rn < 1e-12 or params[n].grad is None: continue
    gr = params[n].grad.numpy().astype(np.float64).ravel()
    p = projection_vectors(n, gr.size) @ gr
    e = float(np.sqrt(np.mean((p - g["fw64_gradproj"][i]) ** 2)) / rn)
    rows.append((e, n, float(g["fw64_ref32err"][i])))
rows.sort(reverse=True)
print("noise", NOISE)
for e, n, r in rows[:8]: print(f"{e:.2e} (ref32 {r:.1e}) {n}")
