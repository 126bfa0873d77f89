"""Shared by the CPU (oracle) and GPU (product) consumers of tests/golden/tdnet_reference_r2.npz."""
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

NPROJ = 16


def projection_vectors(name: str, numel: int) -> np.ndarray:
    rng = np.random.default_rng(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return rng.standard_normal((NPROJ, numel))


def thin(a: np.ndarray) -> np.ndarray:
    return a[:, ::8] if (a.ndim == 3 and a.shape[1] > 200) else (a[:, ::2] if a.ndim == 3 else a)


def rel(a, ref) -> float:
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30))


def check_param_grads(named_grads: dict, gold, tag: str, full_tol: float, proj_tol: float, skip=lambda n: False):
    """Every parameter of the `tag` section of the fixture: tensors stored in full are held to `full_tol` (relative L2); all
    others through their 16 seeded projections: rms_k |<g - g_ref, r_k>| / ||g_ref|| is an unbiased estimate of the
    relative L2 error, held to `proj_tol`. Returns the worst (name, value) for the report."""
    names = [str(n) for n in gold[f"{tag}_names"]]
    worst = ("", 0.0)
    for i, n in enumerate(names):
        if skip(n):
            continue
        ref_norm = float(gold[f"{tag}_gradnorms"][i])
        g = named_grads.get(n)
        if ref_norm < 0:        # the reference leaves this parameter without a gradient (unused q/k/v of the pos_only block)
            assert g is None or float(np.abs(g).max()) == 0.0, n
            continue
        if ref_norm < 1e-12:    # true gradient 0 in fp64 (e.g. a bias feeding straight into BatchNorm, fc_gamma.2.bias under softmax)
            continue
        assert g is not None, n
        key = f"{tag}_grad::" + n
        # yardstick: the reference's OWN fp32 distance from the fp64 truth on this tensor (ReLU flips inside the encoder
        # cannot be masked out of the loss); a tensor passes at max(bar, 3 x that)
        floor = 3.0 * float(gold[f"{tag}_ref32err"][i])
        if key in gold.files:
            e = rel(g, gold[key])
            assert e < max(full_tol, floor), (n, e, floor)
        else:
            p = projection_vectors(n, g.size) @ np.asarray(g, np.float64).ravel()
            e = float(np.sqrt(np.mean((p - gold[f"{tag}_gradproj"][i]) ** 2)) / ref_norm)
            assert e < max(proj_tol, floor), (n, e, floor)
        if e > worst[1]:
            worst = (n, e)
    return worst


def masked_l2(pred, gt, keep):
    """compute_l2_error (model/utils.py:8-11) with kink queries left out of the loss; `keep` is a (B, Q) bool tensor."""
    return (keep.to(pred.dtype) * (pred - gt).pow(2).sum(dim=2) / 2.0).mean()
