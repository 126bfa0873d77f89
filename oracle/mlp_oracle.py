"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

numpy restatement of the plain neural-field MLP of BASELINE.json configs[3] (`3 -> W, n_hidden x (W -> W),
W -> 3`, ReLU between layers) — the per-sample pattern of the reference's decoder
(`model/decoder/crosstransformer_decoder.py:63-69`: nn.Linear stacks with ReLU, `model/decoder/blocks.py:114-142`).
nn.Linear semantics: `y = x @ weight.T + bias` with `weight (out_features, in_features)`.

Parity status: the reference has no module, test or golden vector for this microbenchmark configuration, so the
restatement is pinned against the reference's own building blocks instead: `tests/golden/make_golden_mlp.py` runs a
`torch.nn.Sequential(nn.Linear, nn.ReLU, ...)` stack on seeded inputs and commits inputs/weights/outputs to
`tests/golden/mlp_c4_golden.npz`; `tests/test_oracle_golden.py` holds this file to them.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np


def mlp_forward(x, w_in, b_in, w_h, b_h, w_out, b_out, dtype=np.float64):
    """x (R, Cin); w_in (W, Cin); w_h (n_hidden, W, W); w_out (O, W) in nn.Linear layout -> (R, O)."""
    h = np.asarray(x, dtype=dtype) @ np.asarray(w_in, dtype=dtype).T + np.asarray(b_in, dtype=dtype)
    h = np.maximum(h, 0)
    for l in range(len(w_h)):
        h = h @ np.asarray(w_h[l], dtype=dtype).T + np.asarray(b_h[l], dtype=dtype)
        h = np.maximum(h, 0)
    return h @ np.asarray(w_out, dtype=dtype).T + np.asarray(b_out, dtype=dtype)
