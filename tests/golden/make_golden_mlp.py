"""Mints tests/golden/mlp_c4_golden.npz: a torch.nn.Sequential(nn.Linear, nn.ReLU, ...) stack (the building blocks of the
reference's decoder, model/decoder/blocks.py:114-142) on seeded inputs, fp32 on the CPU. Run once in the authoring
container: `python tests/golden/make_golden_mlp.py`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nsdp_b200 import synth  # noqa: E402

out = {}
for W, L in ((16, 2), (64, 6), (256, 6)):
    w_in, b_in, w_h, b_h, w_out, b_out = synth.mlp_weights(W, L, seed=W)
    layers = [torch.nn.Linear(3, W), torch.nn.ReLU()]
    for _ in range(L):
        layers += [torch.nn.Linear(W, W), torch.nn.ReLU()]
    layers += [torch.nn.Linear(W, 3)]
    net = torch.nn.Sequential(*layers).double()
    lin = [m for m in net if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        lin[0].weight.copy_(torch.from_numpy(w_in)); lin[0].bias.copy_(torch.from_numpy(b_in))
        for l in range(L):
            lin[1 + l].weight.copy_(torch.from_numpy(w_h[l])); lin[1 + l].bias.copy_(torch.from_numpy(b_h[l]))
        lin[-1].weight.copy_(torch.from_numpy(w_out)); lin[-1].bias.copy_(torch.from_numpy(b_out))
        x = (torch.rand(257, 3, generator=torch.Generator().manual_seed(W)) - 0.5).float()
        y64 = net(x.double())
        y32 = net.float()(x)
    out[f"x_{W}"] = x.numpy()
    out[f"y64_{W}"] = y64.numpy()
    out[f"y32_{W}"] = y32.numpy()
    out[f"cfg_{W}"] = np.array([W, L])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mlp_c4_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
