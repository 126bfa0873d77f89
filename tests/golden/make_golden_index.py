"""Golden vectors for the INDEX part of the path (FPS, k-NN, index_points), minted by the LIVE reference's own torch code on the
CPU in the authoring container. Run from the repo root:

    python tests/golden/make_golden_index.py        # -> tests/golden/index_reference.npz

What runs (unmodified, imported from /root/reference/model/utils.py):
  * `farthest_point_sample` (model/utils.py:73-93) — the reference's pure-torch FPS, the alternative its encoder keeps commented
    next to the CUDA call (model/encoder/blocks.py:283-284). Its start index is `torch.randint`; the CUDA kernel starts at 0
    (sampling_gpu.cu:84-86), so `torch.randint` is patched to return 0 for the duration of the call. The clouds have no point
    inside the CUDA kernel's skip radius (|p|^2 <= 1e-3, sampling_gpu.cu:117), where the two would legitimately differ.
  * `square_distance(q, r).argsort()[:, :, :k]` (model/utils.py:39-55, model/encoder/blocks.py:101-102) on clouds whose rows
    have no exact distance tie inside the top k + 1 (asserted below: `argsort` is unstable, ties have no defined order).
  * `index_points` (model/utils.py:58-70).
Inputs ARE stored (float32, a few hundred KB): the consumers must not depend on regenerating them bit-identically.
Consumers: tests/test_index_golden.py (C oracle, no GPU) and tests/test_gpu_index_kernels.py (CUDA kernels through the C ABI).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "index_reference.npz")

from nsdp_b200 import synth  # noqa: E402

spec = importlib.util.spec_from_file_location("nsdp_ref_model_utils", "/root/reference/model/utils.py")
ref_utils = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_utils)


def ref_fps_start0(xyz: torch.Tensor, m: int) -> torch.Tensor:
    real = torch.randint
    torch.randint = lambda lo, hi, size, **kw: torch.zeros(size, dtype=kw.get("dtype", torch.long))
    try:
        return ref_utils.farthest_point_sample(xyz, m)
    finally:
        torch.randint = real


def uniform(B, N, seed):
    return torch.rand(B, N, 3, generator=torch.Generator().manual_seed(seed)) - 0.5


def main():
    gold = {}
    fps_cases = {
        "bumpy_fp16_4096": (synth.surface_cloud(3, 4096, seed=2, fp16_grid=True), 500),    # the model's first down-sampling
        "bumpy_fp32_2048": (synth.surface_cloud(1, 2048, seed=1, fp16_grid=False), 300),
        "bumpy_fp16_500": (synth.surface_cloud(4, 500, seed=4, fp16_grid=True), 100),       # the model's second down-sampling
        "tiny_37": (synth.surface_cloud(2, 37, seed=5, fp16_grid=True), 20),
        "uniform_1000": (uniform(2, 1000, 5), 250),
        "all_points": (uniform(1, 64, 6), 64),                                                  # m == N
    }
    for name, (xyz, m) in fps_cases.items():
        assert float((xyz ** 2).sum(-1).min()) > 1e-3, name      # nothing inside the CUDA kernel's skip radius
        gold[f"fps::{name}::xyz"] = xyz.numpy().astype(np.float32)
        gold[f"fps::{name}::idx"] = ref_fps_start0(xyz, m).numpy().astype(np.int32)

    knn_cases = {
        "self_800_k10": (uniform(2, 800, 3), None, 10),             # TransformerBlock: queries are the cloud itself
        "q300_r1000_k16": (uniform(2, 300, 7), uniform(2, 1000, 8), 16),       # TransformerSetAbstraction: centres vs cloud
        "q500_r100_k7": (uniform(2, 500, 9), uniform(2, 100, 10), 7),          # CrossTransformerBlock: queries vs anchors
        "k_equals_n": (uniform(1, 24, 11), uniform(1, 24, 12), 24),
    }
    for name, (q, r, k) in knn_cases.items():
        r = q if r is None else r
        dist = ref_utils.square_distance(q, r)
        idx = dist.argsort()[:, :, :k]
        srt = dist.sort(dim=-1)[0][:, :, :min(k + 1, r.shape[1])]
        assert bool((srt[:, :, 1:] > srt[:, :, :-1]).all()), f"{name}: a row has a distance tie inside its top k + 1"
        gold[f"knn::{name}::query"] = q.numpy().astype(np.float32)
        gold[f"knn::{name}::ref"] = r.numpy().astype(np.float32)
        gold[f"knn::{name}::idx"] = idx.numpy().astype(np.int32)
        gold[f"knn::{name}::d2"] = torch.gather(dist, 2, idx).numpy().astype(np.float32)

    feats = torch.randn(2, 100, 24, generator=torch.Generator().manual_seed(13))
    idx2 = torch.randint(0, 100, (2, 40), generator=torch.Generator().manual_seed(14))
    idx3 = torch.randint(0, 100, (2, 40, 7), generator=torch.Generator().manual_seed(15))
    gold["index_points::feats"] = feats.numpy()
    gold["index_points::idx2"] = idx2.numpy().astype(np.int32)
    gold["index_points::idx3"] = idx3.numpy().astype(np.int32)
    gold["index_points::out2"] = ref_utils.index_points(feats, idx2).numpy()
    gold["index_points::out3"] = ref_utils.index_points(feats, idx3).numpy()

    np.savez_compressed(OUT, **gold)
    print(f"wrote {OUT}: {len(gold)} arrays, {os.path.getsize(OUT) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
