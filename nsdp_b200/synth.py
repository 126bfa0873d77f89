"""Synthetic clouds, queries and name-keyed weights.

There is no dataset and no checkpoint offline, so tests, smoke() and bench.py all draw their inputs
from here. Everything is a pure function of (name, shape, seed) so that the authoring container, the
GPU box and the committed golden fixtures agree without shipping 18 MB state_dicts.

Value distributions follow the reference pipeline: meshes are PCA-normalised to ~[-0.5, 0.5]
(preprocess/others/process_mesh_local.sh:62-63), samples are stored as float16
(preprocess/generate_dataset_deform4d_surfaceflow.py:74-79) and near-surface queries are perturbed
with sigma 0.1 / 0.02 (preprocess/generate_dataset_deform4d_spaceflow.py:88,106-112).
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch

DEFAULT_MODEL_CFG = {
    "type": "forward",
    "use_normals": False,
    "encoder": "pointransformer",
    "encoder_kwargs": {
        "npoints_per_layer": [5000, 500, 100],
        "nneighbor": 16,
        "nneighbor_reduced": 10,
        "nfinal_transformers": 3,
        "d_transformer": 256,
        "d_reduced": 120,
        "full_SA": True,
    },
    "decoder": "crossatten",
    "decoder_kwargs": {"dim_inp": 256, "dim": 200, "nneigh": 7, "hidden_dim": 128, "out_dim": 3},
}


def make_config(model_type: str = "forward") -> dict:
    """The model block every shipped YAML uses (config/deform4d/forward.yaml:23-41)."""
    import copy
    cfg = copy.deepcopy(DEFAULT_MODEL_CFG)
    cfg["type"] = model_type
    return {"model": cfg, "training": {"optimizer": "Adam", "lr": 5e-4, "lr_step": 200, "lr_decay": 0.1,
                                       "weight_decay": 0.0}}


def make_alt_config() -> dict:
    """A NON-default model block: every knob the reference's blocks honour although no shipped YAML varies it (SURVEY App. A:
    `full_SA: false` -> local final blocks with k = 2 * nneighbor, three down-sampling levels, other widths, `n_blocks`,
    `nneigh`). Exercises the kernels' shape dispatch away from the one configuration the bench runs."""
    cfg = make_config("forward")
    cfg["model"]["encoder_kwargs"] = {"npoints_per_layer": [2000, 400, 128, 48], "nneighbor": 12, "nneighbor_reduced": 8,
                                      "nfinal_transformers": 2, "d_transformer": 128, "d_reduced": 64, "full_SA": False}
    cfg["model"]["decoder_kwargs"] = {"dim_inp": 128, "dim": 96, "nneigh": 5, "hidden_dim": 128, "out_dim": 3, "n_blocks": 3}
    return cfg


def make_ablation_config() -> dict:
    """The reference's ablation blocks behind the same registries: PointNet++-style encoder (max-pool set abstraction)
    and the interpolation decoder (model/encoder/__init__.py:4-7, model/decoder/__init__.py:5-8)."""
    cfg = make_config("forward")
    cfg["model"]["encoder"] = "pointnet++"
    cfg["model"]["encoder_kwargs"] = {"npoints_per_layer": [5000, 500, 100], "nneighbor": 16, "d_transformer": 256,
                                      "nfinal_transformers": 3}
    cfg["model"]["decoder"] = "interp"
    cfg["model"]["decoder_kwargs"] = {"dim_inp": 256, "dim": 200, "hidden_dim": 128, "out_dim": 3}
    return cfg


def _bumpy_sphere(rng: np.random.Generator, n: int, phase: float) -> Tuple[np.ndarray, np.ndarray]:
    """n points on a smooth closed surface of radius ~0.35 and their (approximate) normals."""
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    theta = np.arccos(np.clip(v[:, 2], -1, 1))
    phi = np.arctan2(v[:, 1], v[:, 0])
    r = 0.35 * (1.0 + 0.3 * np.sin(3 * theta + phase) * np.cos(2 * phi))
    return (v * r[:, None]), v


def _smooth_displacement(p: np.ndarray, phase: float, amp: float = 0.12) -> np.ndarray:
    d = np.stack([np.sin(4 * p[:, 1] + phase), np.cos(3 * p[:, 2] - phase), np.sin(5 * p[:, 0] + 2 * phase)], 1)
    return amp * d


def surface_cloud(B: int, N: int, seed: int = 1234, fp16_grid: bool = True) -> torch.Tensor:
    """(B, N, 3) float32 surface samples; fp16_grid rounds through float16 like the stored .npz files
    (exact distance ties then DO occur, which is what the FPS tie-break tests want)."""
    out = np.empty((B, N, 3), np.float32)
    for b in range(B):
        rng = np.random.default_rng(seed * 1000003 + b)
        p, _ = _bumpy_sphere(rng, N, phase=0.37 * b)
        out[b] = p.astype(np.float16).astype(np.float32) if fp16_grid else p.astype(np.float32)
    return torch.from_numpy(out)


def forward_batch(B: int, N: int, Q: int, seed: int = 1234, fp16_grid: bool = True) -> Dict[str, torch.Tensor]:
    """A batch in the layout the datasets produce (dataset/dataset_deform4d_flow.py:217-223):
    surface_samples_inputs (B,N,7) = cat[src, tgt*mask, mask]; space_samples_src/tgt (B,Q,3)."""
    surf = np.empty((B, N, 7), np.float32)
    qs = np.empty((B, Q, 3), np.float32)
    qt = np.empty((B, Q, 3), np.float32)

    def q16(a):
        return a.astype(np.float16).astype(np.float32) if fp16_grid else a.astype(np.float32)

    for b in range(B):
        rng = np.random.default_rng(seed * 1000003 + b)
        p, _ = _bumpy_sphere(rng, N, phase=0.37 * b)
        tgt = p + _smooth_displacement(p, 0.2 * b)
        mask = (p[:, 1] > np.quantile(p[:, 1], 0.7)).astype(np.float64)[:, None]
        surf[b, :, 0:3] = q16(p)
        surf[b, :, 3:6] = q16(tgt) * mask
        surf[b, :, 6:7] = mask
        base, nrm = _bumpy_sphere(rng, Q, phase=0.37 * b)
        sigma = np.where(np.arange(Q) % 2 == 0, 0.1, 0.02)[:, None]
        s = base + nrm * rng.uniform(-1, 1, (Q, 1)) * sigma
        qs[b] = q16(s)
        qt[b] = q16(s + _smooth_displacement(s, 0.2 * b))
    return {"surface_samples_inputs": torch.from_numpy(surf), "space_samples_src": torch.from_numpy(qs),
            "space_samples_tgt": torch.from_numpy(qt)}


def named_tensor(name: str, shape: Iterable[int], seed: int = 0) -> torch.Tensor:
    """Deterministic value for one state_dict entry, a function of (name, shape, seed) only.
    Linear/conv weights ~ U(+-1/sqrt(fan_in)) (torch's default scale); ResnetBlockFC.fc_1.weight
    ~ N(0, 0.05^2) instead of the reference's zero init (decoder/blocks.py:131) so the blocks are
    exercised; BN affine/running stats randomised so eval-mode BN is not the identity."""
    shape = tuple(int(s) for s in shape)
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.long)
    if leaf == "running_mean":
        return torch.randn(shape, generator=g) * 0.1
    if leaf == "running_var":
        return torch.rand(shape, generator=g) + 0.5
    is_bn = any(t in name.split(".")[-2] for t in ("bn", "bnorm")) if "." in name else False
    if is_bn and leaf == "weight":
        return torch.rand(shape, generator=g) * 0.4 + 0.8
    if is_bn and leaf == "bias":
        return torch.randn(shape, generator=g) * 0.1
    if leaf == "weight":
        if name.endswith("fc_1.weight") and ".blocks." in name:
            return torch.randn(shape, generator=g) * 0.05
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
        bound = 1.0 / np.sqrt(max(fan_in, 1))
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    if leaf == "bias":
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.1
    raise ValueError(f"unknown state_dict leaf: {name}")


def named_state_dict(schema: Iterable[Tuple[str, Iterable[int]]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {n: named_tensor(n, s, seed) for n, s in schema}


def mlp_weights(W: int, n_hidden: int, Cin: int = 3, O: int = 3, seed: int = 0):
    """Seeded nn.Linear-style weights (uniform +-1/sqrt(fan_in), as torch's default init) as float32 numpy arrays."""
    rng = np.random.default_rng(seed)

    def lin(o, i):
        b = 1.0 / np.sqrt(i)
        return (rng.uniform(-b, b, size=(o, i)).astype(np.float32), rng.uniform(-b, b, size=(o,)).astype(np.float32))

    w_in, b_in = lin(W, Cin)
    hs = [lin(W, W) for _ in range(n_hidden)]
    w_h = np.stack([h[0] for h in hs]) if n_hidden else np.zeros((0, W, W), np.float32)
    b_h = np.stack([h[1] for h in hs]) if n_hidden else np.zeros((0, W), np.float32)
    # hidden layers are scaled up so that activations neither die nor blow up through 6 ReLU layers
    w_h = (w_h * np.float32(2.4)).astype(np.float32)
    w_out, b_out = lin(O, W)
    return w_in, b_in, w_h, b_h, w_out, b_out
