"""Golden fixtures for the reference's ABLATION blocks (`encoder: pointnet++`, `decoder: interp`;
model/encoder/pointnetplusplus.py, model/decoder/interpolation_decoder.py), produced by running the LIVE reference on
the CPU in the authoring container exactly like make_golden.py does for the shipped configuration:

    python tests/golden/make_golden_ablation.py

Writes tests/golden/ablation_reference.npz and tests/golden/ablation_schema.json.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import OUT, import_reference, schema_of  # noqa: E402

from nsdp_b200 import synth  # noqa: E402


def main():
    torch.set_num_threads(8)
    ref = import_reference()
    cfg = synth.make_ablation_config()
    m, *_ = ref.build_model(cfg)
    schema = schema_of(m)
    with open(os.path.join(OUT, "ablation_schema.json"), "w") as f:
        json.dump(schema, f)
    m.load_state_dict(synth.named_state_dict([(k, s) for k, s in schema], seed=0))
    gold = {}

    # eval-mode forward, single shape (BASELINE.json configs[0] size)
    batch = synth.forward_batch(1, 1024, 2048, seed=1234, fp16_grid=False)
    m.eval()
    with torch.no_grad():
        enc = m.encoder(batch["surface_samples_inputs"])
        out = m(batch["space_samples_src"], batch["surface_samples_inputs"])
    gold["c1_flow"] = out.numpy()
    gold["c1_z"] = enc["z"].numpy()
    gold["c1_anchors"] = enc["anchors"].numpy()
    gold["c1_anchor_feats"] = enc["anchor_feats"].numpy()

    # training step (train-mode BatchNorm), B = 2
    b3 = synth.forward_batch(2, 768, 640, seed=5, fp16_grid=False)
    m.train()
    m.zero_grad()
    q = b3["space_samples_src"].clone().requires_grad_(True)
    pred = m(q, b3["surface_samples_inputs"])
    loss = ref.deformation_networks.compute_l2_error(pred, b3["space_samples_tgt"])
    loss.backward()
    gold["train_loss"] = np.array(loss.item(), np.float64)
    gold["train_pred"] = pred.detach().numpy()
    gold["train_dq"] = q.grad.numpy()
    gold["train_gradnorms"] = np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in m.named_parameters()],
                                       np.float64)
    np.savez_compressed(os.path.join(OUT, "ablation_reference.npz"), **gold)
    print({k: v.shape for k, v in gold.items()})


if __name__ == "__main__":
    main()
